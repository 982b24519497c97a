"""Registration of the B200 decoders under the REFERENCE's own registries.

    import njf_b200.plugin as plugin
    plugin.register_into_reference()          # needs `neural_jacobian_field` importable
    model = neural_jacobian_field.models.model.Model(cfg)   # the reference's Model, now built from B200 decoders

``neural_jacobian_field.models.decoder.{DENSITY_DECODERS, ACTION_DECODERS}`` (models/decoder/__init__.py:11-19) map the
Hydra key ``model.*_decoder.name`` to a class; the factories ``get_density_decoder`` / ``get_action_decoder``
(:30-44) call ``cls(cfg=..., [action_dim=...], encoder_dim=...)``.  The B200 classes take the same arguments, hold
the same parameters under the same names and answer the same per-point methods with the fused kernels, so the
reference's Model / sampler / compositing code runs unchanged on top of them (inference; ``flow_mlp`` keeps its slot
and fails loudly).  For speed use ``njf_b200.Model`` instead: it fuses the sampler and the compositing as well.
"""
from __future__ import annotations

from . import modules

_PREVIOUS = {}


def register_into_reference(strict: bool = True) -> bool:
    """Put the B200 classes into the reference's registries.  Returns False (or raises with strict=True) when the
    reference package is not importable."""
    try:
        import neural_jacobian_field.models.decoder as ref_decoder
    except Exception as ex:  # noqa: BLE001
        if strict:
            raise ImportError("neural_jacobian_field is not importable: nothing to register into") from ex
        return False
    if not _PREVIOUS:
        _PREVIOUS["density"] = dict(ref_decoder.DENSITY_DECODERS)
        _PREVIOUS["action"] = dict(ref_decoder.ACTION_DECODERS)
    ref_decoder.DENSITY_DECODERS.update(modules.DENSITY_DECODERS)
    # "flow_mlp" (ablation decoder without a B200 kernel) keeps the reference's own class
    ref_decoder.ACTION_DECODERS.update({k: v for k, v in modules.ACTION_DECODERS.items() if k != "flow_mlp"})
    return True


def unregister_from_reference() -> None:
    """Restore the reference's own classes."""
    if not _PREVIOUS:
        return
    import neural_jacobian_field.models.decoder as ref_decoder

    ref_decoder.DENSITY_DECODERS.clear()
    ref_decoder.DENSITY_DECODERS.update(_PREVIOUS["density"])
    ref_decoder.ACTION_DECODERS.clear()
    ref_decoder.ACTION_DECODERS.update(_PREVIOUS["action"])
    _PREVIOUS.clear()
