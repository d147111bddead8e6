"""Loading the reference's published / self-trained checkpoints into the drop-in modules.

The reference stores `{'epoch', 'lr', 'random_state', 'optimizer', 'model_pos': state_dict}` (trainval.py:192-199)
and restores with `model_pos.load_state_dict(checkpoint['model_pos'])` (reconstruction.py:238-240,
gen_skes.py:57-58) -- that line works unchanged on the drop-in classes because their `state_dict` key list is
the reference's.  Checkpoints written from an `nn.DataParallel`-wrapped model (trainval.py:56-61) carry a
`module.` prefix on every key; `load_checkpoint` strips it."""
import torch


def strip_module_prefix(state_dict):
    return {(k[len('module.'):] if k.startswith('module.') else k): v for k, v in state_dict.items()}


def load_checkpoint(model_pos, path_or_dict, strict=True):
    """path to a `.bin` written by trainval.py (or the dict itself) -> loads `model_pos`, returns the checkpoint."""
    chk = path_or_dict
    if not isinstance(chk, dict):
        chk = torch.load(chk, map_location=lambda storage, loc: storage)   # reconstruction.py:239
    sd = chk['model_pos'] if 'model_pos' in chk else chk
    model_pos.load_state_dict(strip_module_prefix(sd), strict=strict)
    return chk
