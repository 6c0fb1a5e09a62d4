"""``matplotlib.pyplot`` stand-in: see the package docstring."""
import warnings

import numpy as np

# magma, sampled at 9 equidistant points (approximation of matplotlib's 256-entry table; linear interpolation in between)
_MAGMA = np.array([
    [0.001462, 0.000466, 0.013866], [0.135053, 0.068391, 0.315000], [0.372116, 0.092816, 0.499053],
    [0.594508, 0.175701, 0.501241], [0.828886, 0.262229, 0.430644], [0.973381, 0.461520, 0.361965],
    [0.997341, 0.733545, 0.505167], [0.992440, 0.886330, 0.640580], [0.987053, 0.991438, 0.749504],
])


class _ColourMap:
    def __init__(self, table: np.ndarray, lut: int):
        xs = np.linspace(0.0, 1.0, table.shape[0])
        grid = np.linspace(0.0, 1.0, max(2, int(lut)))
        self._lut = np.stack([np.interp(grid, xs, table[:, c]) for c in range(3)] + [np.ones_like(grid)], axis=-1)

    def __call__(self, values):
        v = np.clip(np.asarray(values, dtype=np.float64), 0.0, 1.0)
        idx = np.minimum((v * self._lut.shape[0]).astype(np.int64), self._lut.shape[0] - 1)
        return self._lut[idx]


def get_cmap(name: str = "magma", lut: int = 256):
    if name != "magma":
        warnings.warn(f"matplotlib stand-in: colour map {name!r} is rendered as 'magma'")
    return _ColourMap(_MAGMA, lut or 256)


class _Anything:
    """Absorbs any attribute access / call (figure, axes, savefig ...): the debug plots are skipped without matplotlib."""

    def __getattr__(self, item):
        return self

    def __call__(self, *args, **kwargs):
        return self

    def __iter__(self):
        return iter(())


_warned = []


def __getattr__(name):
    if not _warned:
        warnings.warn("matplotlib is not installed: plotting calls are no-ops (compat/standins/matplotlib)")
        _warned.append(1)
    return _Anything()
