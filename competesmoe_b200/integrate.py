"""Reference-side binding: put the B200 layers behind the reference's own registries.

The reference discovers MoE layers in two ways: by name through `MOE_REGISTRY` / `get_moe(name)`
(moe_model/model/moe/register.py:4-22, moe_pretrain_model/layers/moe/register.py) and by `isinstance(m, MoeLayer)` /
`isinstance(m, MoE)` (moe_model/train/train.py:1456-1480, llava_trainer.py:1034-1079, llava_arch.py:146-150;
moe_pretrain_model/tasks/simple_task.py:329,388, transformer_lm_mixin.py:263,289).  A drop-in therefore has to be
(a) registered under a name and (b) an instance of the reference's own base class.  `bind_*` builds, at import time of
the reference tree, a subclass of (B200 class, reference base class): method resolution finds the B200 implementation
first, `isinstance` against the reference base holds, and the reference base's `__init__` is never run (the B200
classes initialise `nn.Module` and their mixins explicitly, see multimodal.MoeLayer.__init__ / pretrain.MoE.__init__).

Nothing here touches the hot path; INTEGRATION.md shows the two-line stubs a maintainer adds to the reference tree.
"""
from __future__ import annotations

from typing import Dict, Iterable


def _bind(registry: Dict[str, type], ours: type, ref_base: type, names: Iterable[str], overwrite: bool) -> type:
    bound = type(ours.__name__ + "B200", (ours, ref_base), {"__doc__": ours.__doc__, "__module__": ours.__module__})
    for name in names:
        if name in registry and registry[name] is not bound and not overwrite:
            raise AssertionError(f"MoE name '{name}' is already registered by {registry[name]}; pass overwrite=True "
                                 f"to replace it (the reference's register_moe asserts on conflicts too)")
        registry[name] = bound
    return bound


def bind_multimodal(register_module, moe_module, names=("competesmoe_b200",), overwrite: bool = False) -> type:
    """register_module = moe_model.model.moe.register, moe_module = moe_model.model.moe.moe (already imported).

    `names=("competesmoe",), overwrite=True` replaces the stock layer in place so that existing launch scripts
    (`--moe_name competesmoe`) pick up the B200 kernels; note that moe_model/train/train.py:1487 and
    llava_trainer.py:1037 key extra behaviour on the substring "compete" in the name -- any name containing it works.
    """
    from .multimodal import CompeteSMoE
    return _bind(register_module.MOE_REGISTRY, CompeteSMoE, moe_module.MoeLayer, names, overwrite)


def bind_multimodal_siblings(register_module, moe_module, suffix: str = "_b200", overwrite: bool = False) -> Dict[str, type]:
    """The sibling routers (smoe, smoe_sigmoidgating, xmoe, smoe_perturbed, smoe_share, deepseekv3) under
    `<name><suffix>`; `suffix="", overwrite=True` replaces the stock classes in place."""
    from . import siblings
    from .multimodal import MOE_REGISTRY as ours
    out = {}
    for name in ("smoe", "smoe_sigmoidgating", "xmoe", "smoe_perturbed", "smoe_share", "deepseekv3"):
        out[name] = _bind(register_module.MOE_REGISTRY, ours[name], moe_module.MoeLayer, (name + suffix,), overwrite)
    del siblings
    return out


def bind_pretrain(register_module, moe_module, names=("competesmoe_b200",), overwrite: bool = False) -> type:
    """register_module = layers.moe.register, moe_module = layers.moe.moe of moe_pretrain_model."""
    from .pretrain import CompeteSMoE
    return _bind(register_module.MOE_REGISTRY, CompeteSMoE, moe_module.MoE, names, overwrite)


def bind_pretrain_siblings(register_module, moe_module, suffix: str = "_b200", overwrite: bool = False) -> Dict[str, type]:
    """The pretrain-plugin sibling routers (smoe, smoe_sigmoid, xmoe, smoe_perturbed, deepseekv2, deepseekv3;
    moe_pretrain_model/layers/moe/*.py) under `<name><suffix>`; `suffix="", overwrite=True` replaces them in place."""
    from . import pretrain_siblings
    from .pretrain import MOE_REGISTRY as ours
    out = {}
    for name in ("smoe", "smoe_sigmoid", "xmoe", "smoe_perturbed", "deepseekv2", "deepseekv3"):
        out[name] = _bind(register_module.MOE_REGISTRY, ours[name], moe_module.MoE, (name + suffix,), overwrite)
    del pretrain_siblings
    return out


def bind_cvmm(layers_module) -> None:
    """Replace the Triton op at the package level (moe_pretrain_model/layers/__init__.py:2 re-exports it and the
    11 call sites import `cvmm, cvmm_prepare_sel, cvmm_prepare_sel2, CVMMSel` from there)."""
    from . import cvmm as ours
    for name in ("cvmm", "cvmm_prepare_sel", "cvmm_prepare_sel2", "CVMMSel"):
        setattr(layers_module, name, getattr(ours, name))
