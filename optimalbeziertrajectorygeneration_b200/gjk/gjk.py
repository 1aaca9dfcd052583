"""Drop-in mirror of the reference's ``gjk/gjk.py`` entry point used by the
Bezier distance routines: ``gjkNew(poly1, poly2, maxIter=128, verbose=False)``
(gjk/gjk.py:229-270) -> ``(flag, info)`` with flag 1 / 0 / -1 and, for flag 1,
``info = (point on poly1, point on poly2, distance)``.  Runs one warp per
polygon pair on the GPU (libbezgpu.so, bez_gjk); ``gjk_batch`` takes many pairs
in one launch.  The legacy ``gjkNearest`` family is unused by bezier.py and is
out of scope."""
import ctypes

import numpy as np
import torch

from .. import _capi
from ..engine import F64, _ptr, _require_cuda, _stream


def gjk_batch(polys1, polys2, n1=None, n2=None):
    """polys1 [count, n1max, 3], polys2 [count, n2max, 3]; optional per-item point
    counts.  Returns (flag [count] int32, p1 [count,3], p2 [count,3], dist [count])."""
    _require_cuda()
    dev = torch.device("cuda", torch.cuda.current_device())
    a = torch.as_tensor(np.ascontiguousarray(polys1, dtype=np.float64), device=dev)
    b = torch.as_tensor(np.ascontiguousarray(polys2, dtype=np.float64), device=dev)
    count = int(a.shape[0])
    d_n1 = torch.as_tensor(np.ascontiguousarray(n1, dtype=np.int32), device=dev) if n1 is not None else None
    d_n2 = torch.as_tensor(np.ascontiguousarray(n2, dtype=np.int32), device=dev) if n2 is not None else None
    flag = torch.zeros(count, dtype=torch.int32, device=dev)
    p1 = torch.empty((count, 3), dtype=F64, device=dev)
    p2 = torch.empty((count, 3), dtype=F64, device=dev)
    dist = torch.empty(count, dtype=F64, device=dev)
    ip = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)
    _capi.call("bez_gjk", _ptr(a), _ptr(b), ip(d_n1), ip(d_n2), int(a.shape[1]), int(b.shape[1]), count,
               ip(flag), _ptr(p1), _ptr(p2), _ptr(dist), _stream())
    return flag.cpu().numpy(), p1.cpu().numpy(), p2.cpu().numpy(), dist.cpu().numpy()


def gjkNew(poly1, poly2, maxIter=128, verbose=False):
    """gjk/gjk.py:229-270"""
    if maxIter != 128:
        raise ValueError("the GPU kernel uses the reference's default of 128 iterations")
    poly1 = np.asarray(poly1, dtype=np.float64)
    poly2 = np.asarray(poly2, dtype=np.float64)
    flag, p1, p2, dist = gjk_batch(poly1[None], poly2[None])
    if flag[0] > 0:
        return 1, (p1[0], p2[0], float(dist[0]))
    if flag[0] < 0:
        print('Maximum iterations met')
    return int(flag[0]), ()
