"""game_of_flies_b200: B200-native engine for Game_of_Flies' per-step hot path
(binning -> neighbour forces -> integrate), behind the reference's own call
surface.  Host modules mirror reference main/*.py by name:

    simulation.Simulation            main/simulation.py
    integrator.integrate/setup_bins  main/integrator.py
    acceleration_updater.*           main/acceleration_updater.py
    state_parameters.initialise      main/state_parameters.py
    force_profiles.*                 main/force_profiles.py

Compute is hand-written sm_100a CUDA in csrc/ behind the C ABI of
include/gof_b200.h; there is no CPU fallback.
"""
from ._lib import WRAP_MINIMAGE, WRAP_REFERENCE, GofError  # noqa: F401

__version__ = "0.1.0"
