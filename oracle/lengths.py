"""ORACLE (test infrastructure): numpy restatement of the integer length / alignment code.

  * sequence_mask, generate_path     jyutvoice/utils/model.py:7-11, 29-46
  * make_pad_mask                    jyutvoice/utils/mask.py:232-255
  * duration -> y_lengths            jyutvoice/models/jyutvoice_tts.py:184-187
  * mu_y = attn^T @ mu_x             jyutvoice_tts.py:199-203 (a frame-level gather)

Integer / bool results: bit-exact is the bar.  w_ceil stays float32 exactly as in the reference
(`ceil(w) * length_scale`, scale applied after the ceil; `.long()` truncates).
"""
import numpy as np


def sequence_mask(length, max_length=None):
    length = np.asarray(length)
    if max_length is None:
        max_length = int(length.max())
    x = np.arange(int(max_length), dtype=length.dtype)
    return x[None, :] < length[:, None]


def make_pad_mask(lengths, max_len=0):
    lengths = np.asarray(lengths, dtype=np.int64)
    max_len = max_len if max_len > 0 else int(lengths.max())
    return np.arange(max_len, dtype=np.int64)[None, :] >= lengths[:, None]


def durations(logw, x_mask, length_scale=1.0):
    """logw, x_mask [B,1,Tx] float32 -> (w_ceil [B,1,Tx] float32, y_lengths [B] int64)."""
    logw = np.asarray(logw, dtype=np.float32)
    x_mask = np.asarray(x_mask, dtype=np.float32)
    w = np.exp(logw).astype(np.float32) * x_mask
    w_ceil = (np.ceil(w) * np.float32(length_scale)).astype(np.float32)
    s = w_ceil.sum(axis=(1, 2), dtype=np.float32)
    y_lengths = np.maximum(s, np.float32(1)).astype(np.int64)
    return w_ceil, y_lengths


def generate_path(duration, mask):
    """duration [B,Tx] float32, mask [B,Tx,Ty] -> path [B,Tx,Ty] (0/1, mask dtype)."""
    duration = np.asarray(duration, dtype=np.float32)
    b, t_x, t_y = mask.shape
    cum = np.cumsum(duration, axis=1, dtype=np.float32)
    path = (np.arange(t_y, dtype=np.float32)[None, None, :] < cum[:, :, None]).astype(mask.dtype)
    shifted = np.zeros_like(path)
    shifted[:, 1:] = path[:, :-1]
    return (path - shifted) * mask


def frame_to_token(path):
    """For each mel frame the token index it copies (or -1): the gather form of attn^T @ mu_x."""
    path = np.asarray(path)
    has = path.sum(axis=1) > 0
    idx = path.argmax(axis=1)
    return np.where(has, idx, -1).astype(np.int64)
