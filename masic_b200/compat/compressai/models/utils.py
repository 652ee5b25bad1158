"""compressai.models.utils surface used by MASIC.py:21 (conv, deconv, update_registered_buffers)."""
import torch

from masic_b200.layers import conv, deconv  # noqa: F401


def _update_registered_buffer(module, buffer_name, state_dict_key, state_dict, policy="resize_if_empty",
                              dtype=torch.int):
    new_size = state_dict[state_dict_key].size()
    registered = {n: b for n, b in module.named_buffers()}.get(buffer_name)
    if policy in ("resize_if_empty", "resize"):
        if registered is None:
            raise RuntimeError(f'buffer "{buffer_name}" was not registered')
        if policy == "resize" or registered.numel() == 0:
            registered.resize_(new_size)
    elif policy == "register":
        if registered is not None:
            raise RuntimeError(f'buffer "{buffer_name}" was already registered')
        module.register_buffer(buffer_name, torch.empty(new_size, dtype=dtype).fill_(0))
    else:
        raise ValueError(f'Invalid policy "{policy}"')


def update_registered_buffers(module, module_name, buffer_names, state_dict, policy="resize_if_empty",
                              dtype=torch.int):
    valid = [n for n, _ in module.named_buffers()]
    for name in buffer_names:
        if name not in valid:
            raise ValueError(f'Invalid buffer name "{name}"')
    for name in buffer_names:
        _update_registered_buffer(module, name, f"{module_name}.{name}", state_dict, policy, dtype)
