"""Sequential planning caller (Examples/SequentialSwarm.py:43-70, 176-192).

The reference plans a swarm one vehicle at a time: the constraint of the new vehicle is, for
every already planned trajectory i,
    (new - traj_i).normSquare().elev(R).cpts.min() - maxSep**2
(SequentialSwarm.py:62-67, R = 10), i.e. one scalar per frozen trajectory.  With the new
vehicle as curve 0 the pairs (0, 1), (0, 2), ... are the *first* K pairs of the lexicographic
pair list, so this is the fused pair kernel over the range [0, K) with its in-kernel per-pair
minimum and no row stores (``d_out = NULL``): only the K minima are produced and returned.  x-batches (the FD points of the new vehicle) go through the same
launch.
"""
import ctypes
import os

import numpy as np
import torch

from . import _capi
from . import engine as _engine

F64 = torch.float64


class FrozenSwarm:
    """Already planned trajectories (control points [K, dim, n+1]) resident on the device."""

    def __init__(self, trajectories, elev=10, device=None):
        _engine._require_cuda()
        T = np.ascontiguousarray(np.asarray(trajectories, dtype=np.float64))
        if T.ndim != 3:
            raise ValueError("trajectories must be [K, dim, n+1]")
        self.K, self.dim, n1 = (int(v) for v in T.shape)
        self.n = n1 - 1
        self.elev = int(elev)
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        self.plan = _engine.Plan.get(self.n, self.dim, self.elev, self.device.index)
        self.S = (self.dim * n1 + 1) // 2 * 2
        rows = np.zeros((self.K, self.S))
        rows[:, :self.dim * n1] = T.reshape(self.K, self.dim * n1)
        self.d_rows = torch.as_tensor(rows, device=self.device)

    def separation_minima(self, new_cpts, max_sep, return_device=False):
        """new_cpts [dim, n+1] or [B, dim, n+1] (B candidate curves, e.g. FD points) ->
        [K] or [B, K]: min over the elevated squared-distance control points minus maxSep^2
        (SequentialSwarm.py:temporalSeparationConstraints; note normSquare's dim/2 factor, Q1)."""
        Y = np.asarray(new_cpts, dtype=np.float64)
        single = Y.ndim == 2
        Y = Y[None] if single else Y
        if Y.shape[1:] != (self.dim, self.n + 1):
            raise ValueError("new_cpts must be [dim=%d, n+1=%d]" % (self.dim, self.n + 1))
        B, K, N = Y.shape[0], self.K, self.K + 1
        if K == 0:
            out = np.zeros((B, 0))
            return out[0] if single else out
        cpts = torch.empty((B, N, self.S), dtype=F64, device=self.device)
        cpts[:, 1:] = self.d_rows                                   # frozen trajectories: curves 1..K
        head = np.zeros((B, self.S))
        head[:, :self.dim * (self.n + 1)] = Y.reshape(B, -1)
        cpts[:, 0] = torch.as_tensor(head, device=self.device)      # the new vehicle: curve 0
        minima = torch.empty((B, K), dtype=F64, device=self.device)
        opts = _capi.ReduceOpts()
        opts.itemmin = minima.data_ptr()
        # minima only: the tensor-path kernels skip the K x L rows altogether (reduced epilogue);
        # shapes outside them (degree > 15, L > 128, dim 1) need the rows as scratch
        scratch = None
        if not (self.n <= 15 and self.plan.L <= 128 and self.dim >= 2) or os.environ.get("BEZGPU_FORCE_DFMA") == "1":
            scratch = torch.empty((B, K, self.plan.L), dtype=F64, device=self.device)
        with torch.cuda.device(self.device):
            _capi.call("bez_pair_sepsq_elev_ex", self.plan.handle, _engine._ptr(cpts), B, N, 0, K,
                       float(max_sep) ** 2, _engine._ptr(scratch), ctypes.byref(opts), _engine._stream(self.device))
        if return_device:
            return minima[0] if single else minima
        host = minima.cpu().numpy()
        return host[0] if single else host
