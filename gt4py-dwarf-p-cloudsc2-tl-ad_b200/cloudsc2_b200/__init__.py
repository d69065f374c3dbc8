"""cloudsc2_b200 -- B200-native CLOUDSC2 NL / TL / AD column physics behind the component call
surface of `cloudsc2_gt4py` (module paths mirror the reference: `physics.common.saturation`,
`physics.nonlinear.microphysics`, `physics.tangent_linear.{microphysics,validation}`,
`physics.adjoint.{microphysics,validation}`, `iox`, `setup`)."""
__version__ = "0.1.0"
