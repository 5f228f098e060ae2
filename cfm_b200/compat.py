"""Make ``import torchcfm`` resolve to cfm_b200 so the reference examples run unchanged.

    import cfm_b200.compat; cfm_b200.compat.install_as_torchcfm()
    from torchcfm.conditional_flow_matching import *      # now the B200 implementations
    from torchcfm.models import MLP
    from torchcfm.utils import torch_wrapper
    from torchdyn.core import NeuralODE                    # the lock-step B200 driver

Only the modules of the north_star hot path are provided (optimal_transport,
conditional_flow_matching, models, utils.torch_wrapper, and a ``torchdyn.core.NeuralODE`` stand-in
when torchdyn is not installed); plotting / dataset helpers of the reference are not.
"""
import sys
import types


def install_as_torchcfm(provide_torchdyn=True):
    import cfm_b200
    from . import conditional_flow_matching, models, ode, optimal_transport

    pkg = types.ModuleType("torchcfm")
    pkg.__path__ = []  # mark as package
    for name in conditional_flow_matching.__dict__:
        if not name.startswith("_"):
            setattr(pkg, name, getattr(conditional_flow_matching, name))
    pkg.__version__ = cfm_b200.__version__
    pkg.optimal_transport = optimal_transport
    pkg.conditional_flow_matching = conditional_flow_matching
    mpkg = types.ModuleType("torchcfm.models")
    mpkg.__path__ = []
    mpkg.MLP = models.MLP
    mpkg.models = models
    upkg = types.ModuleType("torchcfm.utils")
    upkg.torch_wrapper = models.torch_wrapper
    pkg.models, pkg.utils = mpkg, upkg
    sys.modules.update({
        "torchcfm": pkg,
        "torchcfm.optimal_transport": optimal_transport,
        "torchcfm.conditional_flow_matching": conditional_flow_matching,
        "torchcfm.models": mpkg,
        "torchcfm.models.models": models,
        "torchcfm.utils": upkg,
    })
    if provide_torchdyn and "torchdyn" not in sys.modules:
        try:
            import torchdyn  # noqa: F401
        except ImportError:
            td = types.ModuleType("torchdyn")
            td.__path__ = []
            core = types.ModuleType("torchdyn.core")
            core.NeuralODE = ode.NeuralODE
            td.core = core
            sys.modules.update({"torchdyn": td, "torchdyn.core": core})
    return pkg
