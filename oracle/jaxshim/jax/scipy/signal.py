import numpy as _np
import scipy.signal as _ss

from .._array import wrap
from ..numpy import _is_taylor


def convolve2d(in1, in2, mode="full", **kw):
    if _is_taylor(in1):
        return in1.map_linear(lambda c: _ss.convolve2d(c, _np.asarray(in2), mode=mode, **kw))
    return wrap(_ss.convolve2d(_np.asarray(in1), _np.asarray(in2), mode=mode, **kw))
