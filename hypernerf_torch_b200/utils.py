"""Checkpoint helpers with the call contract of the reference's utils/__init__.py:66-88 (`extract_model_state_dict`,
`load_ckpt`): Lightning checkpoints keep the model under a module prefix (`nerf.` in train.py:71 / eval.py:137); the
drop-in models use the reference's parameter names, so stripping that prefix is all that is needed."""
from typing import Dict, Iterable

import torch


def _strip(name: str, module: str):
    """`module.rest` -> `rest`; anything else -> None."""
    head = module + "."
    return name[len(head):] if name.startswith(head) else None


def extract_model_state_dict(ckpt_path, model_name: str = 'model', prefixes_to_ignore: Iterable[str] = ()) -> Dict[str, torch.Tensor]:
    """Tensors of sub-module `model_name` from a plain or pytorch-lightning checkpoint, keys relative to that module;
    keys starting with one of `prefixes_to_ignore` (after stripping) are dropped."""
    blob = torch.load(ckpt_path, map_location='cpu')
    tensors = blob.get('state_dict', blob) if isinstance(blob, dict) else blob
    ignore = tuple(prefixes_to_ignore)
    picked = {}
    for full_name, value in tensors.items():
        local = _strip(full_name, model_name)
        if local is None:
            continue
        if ignore and local.startswith(ignore):
            print('ignore', local)
            continue
        picked[local] = value
    return picked


def load_ckpt(model: torch.nn.Module, ckpt_path, model_name: str = 'model', prefixes_to_ignore: Iterable[str] = ()) -> None:
    """Overlay the checkpoint's tensors for `model_name` on the model's current state (missing keys keep their
    values, exactly like the reference); an empty path is a no-op."""
    if not ckpt_path:
        return
    merged = dict(model.state_dict())
    merged.update(extract_model_state_dict(ckpt_path, model_name, prefixes_to_ignore))
    model.load_state_dict(merged)
