"""Stand-in for ``imageio`` (reference trainers.py:194, visualizations/static.py:150,169, render script :112): PNG / GIF output
through Pillow.  ``mimwrite`` to a video container falls back to an animated GIF next to the requested path (no ffmpeg
offline).  Only used when the real package is not installed."""
from pathlib import Path

import numpy as np
from PIL import Image


def _to_image(array) -> Image.Image:
    a = np.asarray(array)
    if a.dtype != np.uint8:
        a = np.clip(a * 255.0 if a.dtype.kind == "f" and a.max() <= 1.0 else a, 0, 255).astype(np.uint8)
    if a.ndim == 3 and a.shape[-1] == 1:
        a = a[..., 0]
    return Image.fromarray(a)


def imwrite(uri, im, **kwargs) -> None:
    _to_image(im).save(str(uri))


imsave = imwrite


def imread(uri, **kwargs):
    return np.asarray(Image.open(str(uri)))


def mimwrite(uri, ims, fps: float = 10.0, **kwargs) -> None:
    frames = [_to_image(f).convert("RGB") for f in ims]
    if not frames:
        return
    path = Path(str(uri))
    if path.suffix.lower() != ".gif":
        path = path.with_suffix(".gif")
    frames[0].save(str(path), save_all=True, append_images=frames[1:], duration=max(1, int(1000.0 / max(fps, 1e-3))), loop=0)


mimsave = mimwrite
