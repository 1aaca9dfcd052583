"""Batches of independent trajectory problems (BASELINE.json configs[4], SURVEY 8(d) "C5").

The reference solves one problem per Python process; its SLSQP loop asks every
constraint callable for the base point and the nvar finite-difference points one after
the other (scipy/optimize/_slsqp_py.py:349-367 -> _numdiff.py:683-712).  Here M problems
that share the model (vehicles, degree, end conditions, bounds) and differ in their
point obstacles and in x are evaluated together: all M x (nvar+1) points of the M FD
sweeps go through the fused kernels in one launch per constraint block, and the
2-point quotients are formed on the device.  Problems are independent, so a multi-GPU
run deals them out in contiguous blocks (sharding.block_range) with no collective.

Only rows that depend on x are produced: the pairs that contain a vehicle (they are a
prefix of the lexicographic pair list because vehicles precede obstacles, Q8); the
obstacle-obstacle rows of the reference's vector are constants.
"""
import numpy as np
import torch

from . import _capi
from . import engine as _engine
from . import optimization as _opt

F64 = torch.float64


class ProblemBatch:
    """``template`` = BezOptimization keyword arguments without ``pointObstacles``;
    ``obstacle_sets`` [M, nObs, dim] = the point obstacles of each problem."""

    def __init__(self, template, obstacle_sets, device=None):
        obstacle_sets = np.ascontiguousarray(obstacle_sets, dtype=np.float64)
        if obstacle_sets.ndim != 3:
            raise ValueError("obstacle_sets must be [M, nObs, dim]")
        self.M, self.nObs = int(obstacle_sets.shape[0]), int(obstacle_sets.shape[1])
        args = dict(template)
        args["pointObstacles"] = [list(o) for o in obstacle_sets[0]]
        self.bezopt = _opt.BezOptimization(**args)
        self.eng = self.bezopt._engine(with_obstacles=True)
        if device is not None and torch.device("cuda", device) != self.eng.device:
            raise ValueError("ProblemBatch lives on the current CUDA device")
        if obstacle_sets.shape[2] != self.eng.dim:
            raise ValueError("obstacles are %d-dimensional, the model is %d-dimensional"
                             % (obstacle_sets.shape[2], self.eng.dim))
        self.d_obst = torch.as_tensor(obstacle_sets, device=self.eng.device)
        self.nvar = self.eng.nvar
        N, nObs = self.eng.N, self.nObs
        self.npairs_x = N * (N - 1) // 2 - nObs * (nObs - 1) // 2     # pairs that contain a vehicle
        self.model = self.bezopt.model

    # ------------------------------------------------------------------
    def fd_points(self, X):
        """host X [M, nvar] -> device (Xp [M, nvar+1, nvar], dx [M, nvar]): the base point and the
        nvar forward points of every problem, h and dx as SciPy chooses them (engine.fd_steps)."""
        X = np.ascontiguousarray(np.atleast_2d(np.asarray(X, dtype=np.float64)))
        if X.shape != (self.M, self.nvar):
            raise ValueError("X must be [%d, %d]" % (self.M, self.nvar))
        h, dx = _engine.ConstraintEngine.fd_steps(X)
        d_x = torch.as_tensor(X, device=self.eng.device)
        d_h = torch.as_tensor(h, device=self.eng.device)
        Xp = d_x[:, None, :].repeat(1, self.nvar + 1, 1)
        idx = torch.arange(self.nvar, device=self.eng.device)
        Xp[:, idx + 1, idx] = d_x + d_h
        return Xp, torch.as_tensor(dx, device=self.eng.device)

    def evaluate(self, d_X, evals_per_problem, elev=None, blocks=("sep", "maxspeed", "angrate")):
        """d_X [M * evals_per_problem, nvar] (device) -> {block: [M * evals_per_problem, m_block]}.
        sep: the npairs_x x-dependent pairs x L; maxspeed: numVeh x L; angrate: numVeh x (4(n+E)+1)."""
        E = _opt._deg_elev() if elev is None else int(elev)
        eng, m = self.eng, self.model
        Q = int(d_X.shape[0])
        cpts, tf = eng.assemble(d_X, E, obst_sets=self.d_obst, evals_per_set=int(evals_per_problem))
        res = {}
        if "sep" in blocks:
            res["sep"] = eng.separation(cpts, E, m["maxSep"], pair_begin=0, npairs=self.npairs_x).view(Q, -1)
        if "maxspeed" in blocks:
            res["maxspeed"] = eng.speed(cpts, tf, E, -1.0, float(m["maxSpeed"]) ** 2).view(Q, -1)
        if "minspeed" in blocks:
            res["minspeed"] = eng.speed(cpts, tf, E, 1.0, -float(m["minSpeed"]) ** 2).view(Q, -1)
        if "angrate" in blocks:
            res["angrate"] = eng.angrate(cpts, tf, E, -1.0, float(m["maxAngRate"]) ** 2).view(Q, -1)
        return res

    def sweep(self, X, elev=None, blocks=("sep", "maxspeed", "angrate")):
        """One finite-difference sweep of every problem: returns {block: (f0 [M, m], JT [M, nvar, m])}
        on the device, JT[p, k] = (f(x_p + h_k e_k) - f(x_p)) / dx_k (SciPy's 2-point formula)."""
        Xp, dx = self.fd_points(X)
        nv1 = self.nvar + 1
        F = self.evaluate(Xp.view(self.M * nv1, self.nvar), nv1, elev, blocks)
        out = {}
        for name, f in F.items():
            mb = int(f.shape[1])
            JT = torch.empty((self.M, self.nvar, mb), dtype=F64, device=f.device)
            _capi.call("bez_fd_quotient_batched", _engine._ptr(f), _engine._ptr(dx), self.M, self.nvar, mb,
                       _engine._ptr(JT), _engine._stream())
            out[name] = (f.view(self.M, nv1, mb)[:, 0], JT)
        return out
