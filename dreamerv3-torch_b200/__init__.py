"""dreamerv3-torch_b200: the DreamerV3 training hot path on B200 (sm_100a) kernels, behind the
module API of ChenFengTsai/dreamerv3-torch.

The directory name carries a dash, so import it with importlib:

    import importlib, sys; sys.path.insert(0, REPO_ROOT)
    dv3 = importlib.import_module("dreamerv3-torch_b200")
    wm = dv3.models.WorldModel(obs_space, act_space, 0, config)

Submodules mirror the reference's files: ``tools`` (lambda_return, DiscDist, OneHotDist,
Optimizer), ``networks`` (RSSM, MLP, encoders / decoders), ``models`` (WorldModel,
ImagBehavior).  ``kernels`` holds the autograd bindings and ``_lib`` the ctypes layer over
libdv3_b200.so; there is no CPU fallback -- calling a hot-path op without the built library
raises ``_lib.Dv3Error``.  ``graphs.TrainStepGraph`` replays a whole WM+AC train step as one CUDA
graph; ``replay`` cuts the reference's replay windows and feeds them to HBM through pinned buffers.
"""
from . import _lib, kernels, tools, networks, models, configs, graphs, replay  # noqa: F401

__all__ = ["_lib", "kernels", "tools", "networks", "models", "configs", "graphs", "replay"]
