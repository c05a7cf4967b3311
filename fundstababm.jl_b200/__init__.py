"""fundstababm.jl_b200 -- B200-native engine for the FundStabABM.jl simulation step.

Host-side mirror of the reference's entry points (``runmodel``, ``Func.initialise``,
``Func.modelrun``, ``Params.default``, ``Types.*``) over the C-ABI library
``csrc/libfsb_b200.so`` (include/fsb.h).  The directory name contains a dot, so import it
through ``__graft_entry__.load_package()`` (module name ``fundstababm_jl_b200``).
"""
from . import params as Params  # noqa: N812  (reference module names)
from . import types as Types  # noqa: N812
from . import func as Func  # noqa: N812
from . import _lib, build, calibrate as _calibrate, engine
from .calibrate import calibrate, grid
from .engine import Engine
from .runmodel import runmodel

__all__ = ["Params", "Types", "Func", "Engine", "runmodel", "calibrate", "grid", "engine", "build", "_lib"]
