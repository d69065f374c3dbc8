"""Look-alikes of the small part of `ifs_physics_common` / `sympl` that the cloudsc2_gt4py
components and harnesses actually use: grid, storage, component base
classes, stencil registry, timing.  Only what the CLOUDSC2 hot path needs is provided."""
