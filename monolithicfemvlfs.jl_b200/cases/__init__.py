"""Case modules mirroring the reference's `src/*.jl` run functions."""
