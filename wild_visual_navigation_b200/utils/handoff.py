"""Weight hand-off between the learner and the per-frame inference.

Reference behaviour restated:
  * writer — ``WvnLearningNode.learning_thread_loop`` (wild_visual_navigation_ros/scripts/wvn_learning_node.py:381-394):
    at most ``load_save_checkpoint_rate`` Hz it saves ``model.state_dict()`` plus one extra key
    ``"confidence_generator" -> cg.get_dict()`` to ``WVN_ROOT_DIR/.tmp_state_dict.pt`` (the old file is removed first);
  * reader — ``WvnFeatureExtractor.load_model`` (wvn_feature_extractor_node.py:407-446): if the file exists and holds
    the model's LAST state-dict key, and that tensor differs from the live one, ``load_state_dict(strict=False)`` and
    copy the confidence generator's ``var / mean / std``.

The on-disk format is kept byte-for-byte compatible (a torch pickle of that dict), so either side can be the
reference's node.  Inside one process ``TraversabilityInference.refresh_weights()`` is the device-to-device
equivalent and needs no file at all.
"""
from __future__ import annotations

import os

import torch

TMP_STATE_DICT_NAME = ".tmp_state_dict.pt"


def write_tmp_state_dict(model: torch.nn.Module, confidence_generator, path: str) -> str:
    """Writer side.  ``path`` is the file (or a directory, in which case ``.tmp_state_dict.pt`` is appended)."""
    if os.path.isdir(path):
        path = os.path.join(path, TMP_STATE_DICT_NAME)
    sd = model.state_dict()
    sd["confidence_generator"] = confidence_generator.get_dict()
    # the reference removes the old file and saves in place (a reader can catch a half-written file); write next to
    # it and rename instead — atomic on POSIX, and the reference's reader sees the same file name and format
    tmp = path + ".writing"
    torch.save(sd, tmp)
    os.replace(tmp, path)
    return path


@torch.no_grad()
def read_tmp_state_dict(model: torch.nn.Module, confidence_generator, path: str) -> bool:
    """Reader side.  Returns True iff new weights were loaded (file present, same model family, weights changed)."""
    if os.path.isdir(path):
        path = os.path.join(path, TMP_STATE_DICT_NAME)
    if not os.path.exists(path):
        return False
    live = model.state_dict()
    device = next(iter(live.values())).device
    try:
        new = torch.load(path, map_location=device, weights_only=False)  # trusted file written by the learner process
    except Exception:  # a reference-side writer saves in place: a half-written file is "no update yet"
        return False
    k = list(live.keys())[-1]
    if k not in new:  # a different model family wrote the file
        return False
    if not (live[k] != new[k].to(device)).any():
        return False
    # in-place copies: the CUDA kernels hold the parameters' storage (SimpleMLP.flat_params), it must not move
    for name, t in live.items():
        if name in new:
            t.copy_(new[name].to(device))
    cg = new.get("confidence_generator")
    if cg is not None and confidence_generator is not None:
        confidence_generator.var.copy_(cg["var"].to(device).reshape(confidence_generator.var.shape))
        confidence_generator.mean.copy_(cg["mean"].to(device).reshape(confidence_generator.mean.shape))
        confidence_generator.std.copy_(cg["std"].to(device).reshape(confidence_generator.std.shape))
    return True
