"""Device side of the input pipeline.

The reference decodes each clip's videos with ffmpeg, converts the frames to float32 NCHW on
the CPU (``datasources/eve_sequences.py:196-211``) and ships 121 MB of floats per B=8, T=30 batch
to the GPU (``core/training.py:256-263``).  Here decoded frames stay uint8 in the decoder's
N x H x W x C layout until they are on the device (30 MB per batch) and one kernel per stream
does the layout change, the scaling and the left/right eye-patch split (``:283-285``), with the
reference's exact float32 rounding.  Video decoding itself (ffmpeg / NVDEC, ``common.py:68-139``)
stays outside this library.
"""
import torch

from . import lib as L


RAW_KEYS = ('eyes_frames', 'screen_frames', 'frames_per_clip')


def _run(frames, x_offset, w_out, scale, bias, frames_per_clip, out=None):
    L.require_cuda(frames, 'preprocess_frames')
    if frames.dtype != torch.uint8:
        raise TypeError('preprocess_frames expects uint8 frames, got %s' % frames.dtype)
    lib = L.load()
    frames = frames.contiguous()
    lead = frames.shape[:-3]
    h, w_in, c = frames.shape[-3:]
    n = 1
    for v in lead:
        n *= int(v)
    steps = int(lead[-1]) if len(lead) >= 2 else 1
    fpc = None
    if frames_per_clip is not None:
        fpc = frames_per_clip.to(device=frames.device, dtype=torch.int32).contiguous()
        assert len(lead) >= 2 and fpc.numel() * steps == n
    if out is None:
        out = torch.empty((*lead, c, h, w_out), dtype=torch.float32, device=frames.device)
    assert out.is_contiguous() and out.dtype == torch.float32 and out.shape == (*lead, c, h, w_out)
    with torch.cuda.device(frames.device):
        L.check(lib.eve_preprocess_frames(L.ptr(frames), n, h, w_in, c, x_offset, w_out, scale, bias,
                                          L.ptr(fpc), steps, L.ptr(out), L.stream_ptr()),
                'eve_preprocess_frames')
    return out


def preprocess_frames(frames, frames_per_clip=None, out=None):
    """eve_sequences.py:196-203: [..., H, W, C] uint8 -> [..., C, H, W] float32 in [-1, 1]."""
    return _run(frames, 0, frames.shape[-2], 2.0 / 255.0, -1.0, frames_per_clip, out)


def preprocess_screen_frames(frames, frames_per_clip=None, out=None):
    """eve_sequences.py:205-211: [..., H, W, C] uint8 -> [..., C, H, W] float32 in [0, 1]."""
    return _run(frames, 0, frames.shape[-2], 1.0 / 255.0, 0.0, frames_per_clip, out)


def preprocess_eye_frames(frames, frames_per_clip=None, out=(None, None)):
    """The "eyes" video holds both patches side by side ([..., H, 2 * ew, C] uint8); returns
    (left_eye_patch, right_eye_patch) as eve_sequences.py:283-285 slices them after
    preprocess_frames: left = columns [ew, 2 ew), right = columns [0, ew)."""
    ew = frames.shape[-2] // 2
    assert frames.shape[-2] == 2 * ew
    left = _run(frames, ew, ew, 2.0 / 255.0, -1.0, frames_per_clip, out[0])
    right = _run(frames, 0, ew, 2.0 / 255.0, -1.0, frames_per_clip, out[1])
    return left, right


def to_device_batch(host_batch, device, non_blocking=True):
    """One training batch as the dataset would hand it over with frames left undecoded-to-float:
    'eyes_frames' [B,T,H,2*ew,3] uint8 (+ 'screen_frames' [B,T,72,128,3] uint8) next to the usual
    small float / bool entries.  Copies everything to `device` and derives left_eye_patch /
    right_eye_patch / screen_frame there."""
    d = {k: v.to(device, non_blocking=non_blocking) for k, v in host_batch.items()}
    fpc = d.pop('frames_per_clip', None)
    if 'eyes_frames' in d:
        d['left_eye_patch'], d['right_eye_patch'] = preprocess_eye_frames(d.pop('eyes_frames'), fpc)
    if 'screen_frames' in d:
        d['screen_frame'] = preprocess_screen_frames(d.pop('screen_frames'), fpc)
    return d
