"""B200-native NeRF ray-render hot path (drop-in for the render/train path of Sahar-E/NeRF-and-DietNeRF).

The directory name follows the project name; import it with
``importlib.import_module("nerf-and-dietnerf_b200")`` or through the alias module ``nerf_b200`` at the repo root.
Sub-modules keep the reference's module names: UtilsCV, UtilsNeuralRadianceField, NeRF, DietNeRF,
ConfigurationKeys.  All compute goes through libnerf_b200.so (hand-written sm_100a CUDA behind a C ABI); there is
no CPU fallback.
"""
from . import ConfigurationKeys, _lib  # noqa: F401
from . import UtilsCV, UtilsNeuralRadianceField, network, optimizers, parallel, poses, vit  # noqa: F401
from . import NeRF as _nerf_module, DietNeRF as _dietnerf_module  # noqa: F401
from . import ExecutionRun as _execution_run_module, UtilsFiles, UtilsVideo, h5weights  # noqa: F401
from ._lib import LIB_PATH, NerfLibraryError, NetCfg, load  # noqa: F401
from .network import NerfMLP  # noqa: F401
from .optimizers import Adam  # noqa: F401

NeRFModel = _nerf_module.NeRF
DietNeRFModel = _dietnerf_module.DietNeRF
ExecutionRun = _execution_run_module.ExecutionRun
