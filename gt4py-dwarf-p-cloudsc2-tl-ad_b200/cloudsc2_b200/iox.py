"""Parameter models and HDF5 scalar reader.

Mirrors `cloudsc2_gt4py/iox.py` of the reference:
  * `YoethfParams` (iox.py:25-45), `YomcstParams` (:48-57), `YrecldpParams` (:60-182),
    `YrephliParams` (:185-201), `YrnclParams` (:204-205), `YrphncParams` (:208-209):
    same attribute names, `.dict()` like the pydantic-v1 models, attributes mutable in place
    (the Taylor harness does `yrncl_params.LREGCL = False`, tangent_linear/validation.py:85).
  * `HDF5Operator` (:212-244): `get_nlev/get_nlon/get_timestep/get_*_params`, reading scalar
    datasets named after the attributes (prefix `YRECLDP_` / `YREPHLI_` where the reference
    uses one) through the pure-Python reader in `h5lite.py` (no h5py in this image).

The reference reads every constant from `data/input.h5`, which is not shipped.  The
`ifs_defaults()` table below holds the IFS standard values (from the IFS documentation /
upstream dwarf, from memory -- see DESIGN.md "Constants"); they are overridden by the file
values whenever an `input.h5` is available.
"""
from __future__ import annotations

from datetime import timedelta
from typing import Any, Callable, Dict, Optional


class _Params:
    """Minimal stand-in for a pydantic-v1 BaseModel: keyword construction, `.dict()`,
    attribute mutation.  Unknown keywords are kept (the reference models carry ~140 unused
    YRECLDP members which still travel into the externals dict)."""

    _fields: Dict[str, Any] = {}

    def __init__(self, **kwargs: Any) -> None:
        for name, default in self._fields.items():
            if name in kwargs:
                value = kwargs.pop(name)
            elif default is not _REQUIRED:
                value = default
            else:
                raise TypeError(f"{type(self).__name__}: missing parameter {name!r}")
            if isinstance(default, bool) or isinstance(value, bool) or type(value).__name__ == "bool_":
                value = bool(value)
            elif hasattr(value, "dtype") and value.dtype.kind in "iu":
                value = int(value)
            elif not isinstance(value, int):
                value = float(value)
            object.__setattr__(self, name, value)
        for name, value in kwargs.items():  # extra members are preserved, not validated
            object.__setattr__(self, name, value)

    def dict(self) -> Dict[str, Any]:
        return dict(self.__dict__)

    def copy(self):
        return type(self)(**self.dict())

    def __repr__(self) -> str:
        return f"{type(self).__name__}({', '.join(f'{k}={v!r}' for k, v in self.__dict__.items())})"


class _Required:
    def __repr__(self) -> str:
        return "<required>"


_REQUIRED = _Required()


class YoethfParams(_Params):
    """reference iox.py:25-45"""

    _fields = {
        n: _REQUIRED
        for n in (
            "R2ES R3IES R3LES R4IES R4LES R5ALSCP R5ALVCP R5IES R5LES RALFDCP RALSDCP RALVDCP "
            "RKOOP1 RKOOP2 RTICE RTICECU RTWAT RTWAT_RTICECU_R RTWAT_RTICE_R"
        ).split()
    }
    _fields["RVTMP2"] = 0.0


class YomcstParams(_Params):
    """reference iox.py:48-57"""

    _fields = {n: _REQUIRED for n in "RCPD RD RETV RG RLMLT RLSTT RLVTT RTT RV".split()}


class YrecldpParams(_Params):
    """reference iox.py:60-182.  Only RCLCRIT, RKCONV, RLMIN, RPECONS are read by the
    stencils (nonlinear/_stencils/cloudsc2.py:61-91); every other member of the reference
    model is accepted as an extra keyword and carried along untouched."""

    _fields = {n: _REQUIRED for n in "RCLCRIT RKCONV RLMIN RPECONS".split()}


class YrephliParams(_Params):
    """reference iox.py:185-201.  Only RLPTRC is read by the stencils."""

    _fields = {"RLPTRC": _REQUIRED, "LPHYLIN": True}


class YrnclParams(_Params):
    """reference iox.py:204-205"""

    _fields = {"LREGCL": True}


class YrphncParams(_Params):
    """reference iox.py:208-209"""

    _fields = {"LEVAPLS2": False}


def ifs_defaults() -> Dict[str, _Params]:
    """IFS standard constants (SUCST / SUETHF / SUCLDP / SUPHLI values).  UNVERIFIED against
    `input.h5` (not shipped); RLSTT and RLVTT are confirmed by the golden outputs
    (PFHPSN = -RLSTT * PFPLSN to 1 ulp in data/reference_double.h5)."""
    RKBOL, RNAVO = 1.380658e-23, 6.0221367e23
    R = RNAVO * RKBOL
    RMD, RMV = 28.9644, 18.0153
    RD = 1000.0 * R / RMD
    RV = 1000.0 * R / RMV
    RCPD = 3.5 * RD
    RETV = RV / RD - 1.0
    RG = 9.80665
    RTT = 273.16
    RLVTT, RLSTT = 2.5008e6, 2.8345e6
    RLMLT = RLSTT - RLVTT
    R3LES, R3IES, R4LES, R4IES = 17.502, 22.587, 32.19, -0.7
    R5LES = R3LES * (RTT - R4LES)
    R5IES = R3IES * (RTT - R4IES)
    RTWAT = RTT
    RTICE = RTT - 23.0
    RTICECU = RTT - 23.0
    yomcst = YomcstParams(RCPD=RCPD, RD=RD, RETV=RETV, RG=RG, RLMLT=RLMLT, RLSTT=RLSTT, RLVTT=RLVTT, RTT=RTT, RV=RV)
    yoethf = YoethfParams(
        R2ES=611.21 * RD / RV, R3IES=R3IES, R3LES=R3LES, R4IES=R4IES, R4LES=R4LES,
        R5ALSCP=R5IES * RLSTT / RCPD, R5ALVCP=R5LES * RLVTT / RCPD, R5IES=R5IES, R5LES=R5LES,
        RALFDCP=RLMLT / RCPD, RALSDCP=RLSTT / RCPD, RALVDCP=RLVTT / RCPD,
        RKOOP1=2.583, RKOOP2=0.48116e-2, RTICE=RTICE, RTICECU=RTICECU, RTWAT=RTWAT,
        RTWAT_RTICECU_R=1.0 / (RTWAT - RTICECU), RTWAT_RTICE_R=1.0 / (RTWAT - RTICE), RVTMP2=0.0,
    )
    yrecldp = YrecldpParams(RCLCRIT=0.4e-3, RKCONV=1.0 / 6000.0, RLMIN=1.0e-8, RPECONS=5.44e-4 / RG)
    yrephli = YrephliParams(RLPTRC=266.42345, LPHYLIN=True)
    return {
        "yoethf": yoethf, "yomcst": yomcst, "yrecldp": yrecldp, "yrephli": yrephli,
        "yrncl": YrnclParams(), "yrphnc": YrphncParams(),
    }


DEFAULT_TIMESTEP = timedelta(seconds=3600.0)  # PTSPHY of the CLOUDSC dwarfs' input


class HDF5Operator:
    """reference iox.py:212-244 on top of `h5lite.File` (contiguous datasets only)."""

    def __init__(self, filename: str, gt4py_config: Any = None) -> None:
        from .h5lite import File

        self.f = File(filename)
        self.gt4py_config = gt4py_config

    def get_nlev(self) -> int:
        return int(self.f["KLEV"][0])

    def get_nlon(self) -> int:
        return int(self.f["KLON"][0])

    def get_timestep(self) -> timedelta:
        return timedelta(seconds=float(self.f["PTSPHY"][0]) if "PTSPHY" in self.f else 0.0)

    def get_params(self, cls, get_param_name: Optional[Callable[[str], str]] = None):
        get_param_name = get_param_name or (lambda name: name)
        kwargs = {}
        for name, default in cls._fields.items():
            h5_name = get_param_name(name)
            if h5_name in self.f:
                kwargs[name] = self.f[h5_name][0]
            elif default is _REQUIRED:
                raise KeyError(f"{h5_name} not found in {self.f.filename}")
        return cls(**kwargs)

    def get_yoethf_params(self) -> YoethfParams:
        return self.get_params(YoethfParams)

    def get_yomcst_params(self) -> YomcstParams:
        return self.get_params(YomcstParams)

    def get_yrecldp_params(self) -> YrecldpParams:
        return self.get_params(YrecldpParams, lambda n: "YRECLDP_" + n)

    def get_yrephli_params(self) -> YrephliParams:
        return self.get_params(YrephliParams, lambda n: "YREPHLI_" + n)

    def get_yrncl_params(self) -> YrnclParams:
        return self.get_params(YrnclParams)

    def get_yrphnc_params(self) -> YrphncParams:
        return self.get_params(YrphncParams)
