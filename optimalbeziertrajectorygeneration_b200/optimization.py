"""Drop-in mirror of the reference's ``optimization.py`` whose constraint and
cost callables run on the B200 (libbezgpu.so) instead of numpy/numba.

Same names, argument meaning and error behaviour as the reference:
``BezOptimization(numVeh, dimension, degree, minimizeGoal, maxSep, minSpeed,
maxSpeed, maxAngRate, initPoints, finalPoints, initSpeeds, finalSpeeds,
initAngs, finalAngs, tf, pointObstacles, shapeObstacles)`` with *properties
returning closures* ``f(x: float64[nvar]) -> float64[m]`` that are handed to
``scipy.optimize.minimize(method='SLSQP')`` (optimization.py:20-187), the module
global ``DEG_ELEV`` that is re-read at call time (optimization.py:17),
``generateGuess`` and ``reshapeVector``.

Additive API (not in the reference): ``*_jac`` closures returning the
finite-difference Jacobian SciPy would have formed with nvar+1 calls, computed
in one batched launch; ``evaluate_batch`` for device-resident batched
evaluation.
"""
import numpy as np
import torch

from . import engine as _engine

DEG_ELEV = 0


def _deg_elev():
    # re-read the module global at call time, like the reference (Q9)
    return int(globals()['DEG_ELEV'])


class BezOptimization:
    def __init__(self,
                 numVeh=1,
                 dimension=1,
                 degree=5,
                 minimizeGoal='Euclidean',
                 maxSep=0.9,
                 minSpeed=0,
                 maxSpeed=1e6,
                 maxAngRate=1e6,
                 initPoints=None,
                 finalPoints=None,
                 initSpeeds=None,
                 finalSpeeds=None,
                 initAngs=None,
                 finalAngs=None,
                 tf=1.0,
                 pointObstacles=None,
                 shapeObstacles=None,
                 device=None):
        self.pointObstacles = pointObstacles
        self.shapeObstacles = shapeObstacles

        self._numCols = degree + 1
        if initPoints is not None:
            self._numCols -= 2
        if initSpeeds is not None:
            self._numCols -= 2

        # optimization.py:49-63
        self.model = {'numVeh': numVeh,
                      'dim': dimension,
                      'deg': degree,
                      'minGoal': minimizeGoal,
                      'maxSep': maxSep,
                      'minSpeed': minSpeed,
                      'maxSpeed': maxSpeed,
                      'maxAngRate': maxAngRate,
                      'initPoints': np.atleast_2d(initPoints),
                      'finalPoints': np.atleast_2d(finalPoints),
                      'initSpeeds': np.atleast_1d(initSpeeds),
                      'finalSpeeds': np.atleast_1d(finalSpeeds),
                      'initAngs': np.atleast_1d(initAngs),
                      'finalAngs': np.atleast_1d(finalAngs),
                      'tf': tf}
        self._device = device
        # additive: when True the closures return views of the pinned staging
        # buffer (valid until the next call of the same closure) instead of copies
        self.zero_copy_results = False
        self._engines = {}

    # ------------------------------------------------------------------
    def _engine(self, with_obstacles):
        """Device state; built lazily so constructing a model needs no GPU."""
        key = bool(with_obstacles) and self.pointObstacles is not None
        eng = self._engines.get(key)
        if eng is None:
            eng = _engine.ConstraintEngine(self.model, self.pointObstacles if key else None,
                                           device=self._device)
            self._engines[key] = eng
        return eng

    @property
    def nvar(self):
        return (self.model['numVeh'] * self.model['dim'] * self._numCols +
                (1 if self.model['minGoal'].lower() == 'timeopt' else 0))

    # ------------------------------------------------------------------
    @property
    def objectiveFunction(self):
        minGoal = self.model['minGoal'].lower()

        objectivesDict = {'euclidean': self.euclideanObjective,
                          'timeopt': lambda x: x[-1],
                          'accel': self.accelObjective,
                          'jerk': self.jerkObjective,
                          }
        try:
            return objectivesDict[minGoal]
        except KeyError:
            err = ('The provided minimize goal, {}, is not a valid goal. '
                   'The available minimize goals are:\n{}'
                   ).format(minGoal, objectivesDict.keys())
            raise ValueError(err)

    # ------------------------------------------------------------------
    @property
    def temporalSeparationConstraints(self):
        """optimization.py:83-107 -> _temporalSeparationConstraints (:311-346)."""
        def wrapper(x):
            eng = self._engine(with_obstacles=True)
            if eng.N <= 1:
                return None                       # optimization.py:345-346
            E = _deg_elev()
            cpts, _ = eng.assemble(eng.upload(x), E)
            out = eng.separation(cpts, E, self.model['maxSep'])
            return eng.download(out, copy=not self.zero_copy_results).reshape(-1)
        return wrapper

    def spatialSeparationConstraints(self, x):
        """optimization.py:109-133: all pairs among the vehicles and the shape
        obstacles (Bezier objects), minDist - maxSep; like the reference the
        (alpha, t1, t2) tuple is kept, so the result is [npairs, 3].
        All pairs run in one launch (one warp per pair)."""
        from . import bezier as _bez
        numVeh = self.model['numVeh']
        dim = self.model['dim']
        maxSep = self.model['maxSep']
        y = self.reshapeVector(x)
        curves = [np.ascontiguousarray(y[i * dim:(i + 1) * dim, :]) for i in range(numVeh)]
        for obstacle in (self.shapeObstacles or []):
            curves.append(np.ascontiguousarray(obstacle.cpts, dtype=np.float64))
        n = len(curves)
        shapes = {c.shape for c in curves}
        if len(shapes) != 1:
            raise ValueError('all vehicles and shape obstacles must share dimension and degree')
        A = np.stack([curves[i] for i in range(n) for j in range(i + 1, n)])
        Bc = np.stack([curves[j] for i in range(n) for j in range(i + 1, n)])
        out, status = _bez.min_dist_batch(A, Bc, max_nodes=1 << 18)
        self.last_status = status
        return out - maxSep

    @property
    def minSpeedConstraints(self):
        """optimization.py:135-151 -> _minSpeedConstraints (:349-384)."""
        def wrapper(x):
            eng = self._engine(with_obstacles=False)
            E = _deg_elev()
            cpts, tf = eng.assemble(eng.upload(x), E)
            out = eng.speed(cpts, tf, E, 1.0, -float(self.model['minSpeed']) ** 2)
            return eng.download(out, copy=not self.zero_copy_results).reshape(-1)
        return wrapper

    @property
    def maxSpeedConstraints(self):
        """optimization.py:153-169 -> _maxSpeedConstraints (:387-422)."""
        def wrapper(x):
            eng = self._engine(with_obstacles=False)
            E = _deg_elev()
            cpts, tf = eng.assemble(eng.upload(x), E)
            out = eng.speed(cpts, tf, E, -1.0, float(self.model['maxSpeed']) ** 2)
            return eng.download(out, copy=not self.zero_copy_results).reshape(-1)
        return wrapper

    @property
    def maxAngularRateConstraints(self):
        """optimization.py:171-187 -> _maxAngularRateConstraints (:425-459)."""
        def wrapper(x):
            if self.model['dim'] != 2:                      # optimization.py:590-593
                msg = ('The input curve must be two dimensional,\n'
                       'instead it is {} dimensional'.format(self.model['dim']))
                raise ValueError(msg)
            eng = self._engine(with_obstacles=False)
            E = _deg_elev()
            cpts, tf = eng.assemble(eng.upload(x), E)
            out = eng.angrate(cpts, tf, E, -1.0, float(self.model['maxAngRate']) ** 2)
            return eng.download(out, copy=not self.zero_copy_results).reshape(-1)
        return wrapper

    # ------------------------------------------------------------------
    # additive: Jacobians of the closures above, J[m, nvar], equal to what SLSQP
    # would form by nvar+1 calls (SciPy's '2-point' rule and step), computed in one
    # batched launch.  Hand them to SciPy as {'type': 'ineq', 'fun': f, 'jac': f_jac}.
    @property
    def temporalSeparationConstraints_jac(self):
        def wrapper(x):
            eng = self._engine(with_obstacles=True)
            if eng.N <= 1:
                return None
            JT = eng.jac_separation(x, _deg_elev(), dense=True)
            return eng.download(JT, key="jac", copy=not self.zero_copy_results).T
        return wrapper

    @property
    def maxSpeedConstraints_jac(self):
        def wrapper(x):
            eng = self._engine(with_obstacles=False)
            JT = eng.jac_speed(x, _deg_elev(), -1.0, dense=True)
            return eng.download(JT, key="jac", copy=not self.zero_copy_results).T
        return wrapper

    @property
    def minSpeedConstraints_jac(self):
        def wrapper(x):
            eng = self._engine(with_obstacles=False)
            JT = eng.jac_speed(x, _deg_elev(), 1.0, dense=True)
            return eng.download(JT, key="jac", copy=not self.zero_copy_results).T
        return wrapper

    @property
    def maxAngularRateConstraints_jac(self):
        """Batched literal FD (no closed form for the ratio): same noise floor
        as SciPy differencing the reference (~1e-7 relative)."""
        def wrapper(x):
            eng = self._engine(with_obstacles=False)
            JT = eng.jac_angrate(x, _deg_elev(), -1.0, float(self.model['maxAngRate']) ** 2)
            return eng.download(JT, key="jac", copy=not self.zero_copy_results).T
        return wrapper

    # ------------------------------------------------------------------
    def generateGuess(self, std=0, seed=None):
        """optimization.py:189-240 (host logic; uses numpy's global RNG like
        the reference so seeded guesses are identical)."""
        dim = self.model['dim']
        deg = self.model['deg']
        numVeh = self.model['numVeh']
        tf = self.model['tf']
        initPoints = self.model['initPoints']
        finalPoints = self.model['finalPoints']
        initSpeeds = self.model['initSpeeds']
        finalSpeeds = self.model['finalSpeeds']
        initAngs = self.model['initAngs']
        finalAngs = self.model['finalAngs']

        np.random.seed(seed)
        xGuess = []
        for i in range(numVeh):
            for j in range(dim):
                if initSpeeds[0] is None:
                    line = np.linspace(initPoints[i, j], finalPoints[i, j], deg + 1)
                    line += np.random.randn(deg + 1) * std
                else:
                    if dim != 2:
                        err = ('The dimension must be 2 for initial and final '
                               'speeds and angles.')
                        raise ValueError(err)
                    initMag = initSpeeds[i] * tf / deg
                    finalMag = finalSpeeds[i] * tf / deg
                    if j % 2 == 0:
                        initPt = initPoints[i, j] + initMag * np.cos(initAngs[i])
                        finalPt = finalPoints[i, j] - finalMag * np.cos(finalAngs[i])
                    else:
                        initPt = initPoints[i, j] + initMag * np.sin(initAngs[i])
                        finalPt = finalPoints[i, j] - finalMag * np.sin(finalAngs[i])
                    line = np.linspace(initPt, finalPt, deg + 1 - 2)
                    line += np.random.randn(deg + 1 - 2) * std
                xGuess.append(line[1:-1])
        if self.model['minGoal'].lower() == 'timeopt':
            xGuess.append([tf])
        return np.concatenate(xGuess)

    def reshapeVector(self, x):
        """optimization.py:242-285, evaluated by the device assemble kernel
        (the same one every constraint closure uses) and copied back as the
        reference's [numVeh*dim, deg+1] matrix."""
        eng = self._engine(with_obstacles=False)
        cpts, _ = eng.assemble(eng.upload(x), 0)             # [1, numVeh, S]
        rows = self.model['dim'] * (self.model['deg'] + 1)
        y = eng.download(cpts[0, :, :rows].contiguous())
        return y.reshape(self.model['numVeh'] * self.model['dim'], self.model['deg'] + 1)

    # ------------------------------------------------------------------
    # additive, batched API
    def evaluate_batch(self, X, which=('sep', 'maxspeed'), elev=None):
        """Evaluates the named constraint blocks for every row of X [B, nvar]
        in one pass; returns device tensors {name: [B, m_name]}."""
        E = _deg_elev() if elev is None else int(elev)
        res = {}
        if 'sep' in which:
            eng = self._engine(with_obstacles=True)
            d_x = X if isinstance(X, torch.Tensor) else eng.upload(X)
            cpts, _ = eng.assemble(d_x, E)
            res['sep'] = eng.separation(cpts, E, self.model['maxSep']).reshape(d_x.shape[0], -1)
        eng = self._engine(with_obstacles=False)
        if 'maxspeed' in which or 'minspeed' in which:
            d_x = X if isinstance(X, torch.Tensor) else eng.upload(X)
            cpts, tf = eng.assemble(d_x, E)
            if 'maxspeed' in which:
                res['maxspeed'] = eng.speed(cpts, tf, E, -1.0, float(self.model['maxSpeed']) ** 2
                                            ).reshape(d_x.shape[0], -1)
            if 'minspeed' in which:
                res['minspeed'] = eng.speed(cpts, tf, E, 1.0, -float(self.model['minSpeed']) ** 2
                                            ).reshape(d_x.shape[0], -1)
        return res

    def evaluate_reduced(self, X, elev=None):
        """Additive API for swarms whose full constraint vector is consumed on the
        device (508 MB per x at N=1024): host X [B, nvar] -> host arrays
        ``pairmin`` [B, P] (min over each pair's elevated separation values; its
        sign is the active-pair flag) and ``maxspeed`` [B, numVeh*L].  Every
        separation row is still materialised in HBM and stays available as the
        device tensor ``self.workspace['sep']`` until the next call."""
        E = _deg_elev() if elev is None else int(elev)
        eng = self._engine(with_obstacles=True)
        X = np.atleast_2d(np.asarray(X, dtype=np.float64))
        B = X.shape[0]
        P = _engine.num_pairs(eng.N)
        L = 2 * self.model['deg'] + E + 1
        ws = getattr(self, 'workspace', None)
        if ws is None or ws.get('key') != (B, E):
            ws = {'key': (B, E),
                  'sep': torch.empty((B, P, L), dtype=torch.float64, device=eng.device),
                  'pairmin': torch.empty((B, P), dtype=torch.float64, device=eng.device),
                  'maxspeed': torch.empty((B, self.model['numVeh'], L), dtype=torch.float64,
                                          device=eng.device)}
            self.workspace = ws
        cpts, tf = eng.assemble(eng.upload(X), E)
        eng.separation(cpts, E, self.model['maxSep'], out=ws['sep'], pairmin=ws['pairmin'])
        eng.speed(cpts, tf, E, -1.0, float(self.model['maxSpeed']) ** 2, nveh=self.model['numVeh'],
                  out=ws['maxspeed'])
        # two async copies into pinned staging, one synchronisation
        pm = eng._pinned_buf('pairmin', ws['pairmin'].numel())
        sp = eng._pinned_buf('maxspeed', ws['maxspeed'].numel())
        pm.copy_(ws['pairmin'].view(-1), non_blocking=True)
        sp.copy_(ws['maxspeed'].view(-1), non_blocking=True)
        torch.cuda.current_stream().synchronize()
        pmh, sph = pm.numpy().reshape(B, P), sp.numpy().reshape(B, -1)
        if not self.zero_copy_results:
            pmh, sph = pmh.copy(), sph.copy()
        return {'pairmin': pmh, 'maxspeed': sph}

    def evaluate_sweep(self, X, elev=None, chunk=4, out=None):
        """Pipelined form of :meth:`evaluate_reduced` for the nvar+1 points of a finite
        difference sweep (SciPy asks for them one by one, `_numdiff.py:683-712`; here the
        caller hands all of them over): host X [M, nvar] is processed in chunks of
        ``chunk`` rows with two device workspaces, so that the pinned host->device copy
        and the kernels of chunk k+1 overlap the device->host copy of chunk k (separate
        copy stream, events, no host synchronisation inside the loop).  Returns host
        arrays ``pairmin`` [M, P] and ``maxspeed`` [M, numVeh*L] (pinned; reused by the
        next call unless ``out`` supplies the destination dict of pinned tensors)."""
        E = _deg_elev() if elev is None else int(elev)
        eng = self._engine(with_obstacles=True)
        X = np.ascontiguousarray(np.atleast_2d(np.asarray(X, dtype=np.float64)))
        M = X.shape[0]
        if X.shape[1] != eng.nvar:
            raise ValueError("x has %d entries, the model expects %d" % (X.shape[1], eng.nvar))
        P = _engine.num_pairs(eng.N)
        L = 2 * self.model['deg'] + E + 1
        nv = self.model['numVeh']
        chunk = max(1, min(int(chunk), M))
        sw = getattr(self, '_sweep_ws', None)
        if sw is None or sw['key'] != (chunk, E):
            def wsset():
                return {'x': torch.empty((chunk, eng.nvar), dtype=torch.float64, device=eng.device),
                        'sep': torch.empty((chunk, P, L), dtype=torch.float64, device=eng.device),
                        'pairmin': torch.empty((chunk, P), dtype=torch.float64, device=eng.device),
                        'maxspeed': torch.empty((chunk, nv, L), dtype=torch.float64, device=eng.device),
                        'done': None}
            sw = {'key': (chunk, E), 'sets': [wsset(), wsset()], 'copy_stream': torch.cuda.Stream(device=eng.device),
                  'copy_stream2': torch.cuda.Stream(device=eng.device),
                  'upload_stream': torch.cuda.Stream(device=eng.device)}
            self._sweep_ws = sw
        if out is None:
            out = {'pairmin': eng._pinned_buf('sweep_pairmin', M * P).view(M, P),
                   'maxspeed': eng._pinned_buf('sweep_maxspeed', M * nv * L).view(M, nv * L)}
        xs = eng._pinned_buf('sweep_x', X.size).view(M, eng.nvar)
        xs_np = xs.numpy()                           # filled chunk by chunk, just ahead of each upload
        main, side, side2 = torch.cuda.current_stream(), sw['copy_stream'], sw['copy_stream2']
        up = sw['upload_stream']
        max_speed2 = float(self.model['maxSpeed']) ** 2

        def upload(k, after):
            # x of chunk k goes up one step ahead on its own stream, so that it never queues
            # behind the result copies of the previous chunk on a shared copy engine
            lo_ = k * chunk
            hi_ = min(M, lo_ + chunk)
            xs_np[lo_:hi_] = X[lo_:hi_]
            with torch.cuda.stream(up):
                if after is not None:
                    up.wait_event(after)             # kernels that last read this set's x
                sw['sets'][k & 1]['x'][:hi_ - lo_].copy_(xs[lo_:hi_], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(up)
            return ev

        nchunks = (M + chunk - 1) // chunk
        up.wait_stream(main)
        x_ready = upload(0, None)
        prev_ready = None
        for k, lo in enumerate(range(0, M, chunk)):
            hi = min(M, lo + chunk)
            b = hi - lo
            ws = sw['sets'][k & 1]
            if ws['done'] is not None:
                main.wait_event(ws['done'])          # D2H of the chunk that used this set two steps ago
            main.wait_event(x_ready)
            if k + 1 < nchunks:
                x_ready = upload(k + 1, prev_ready)
            cpts, tf = eng.assemble(ws['x'][:b], E)
            eng.separation(cpts, E, self.model['maxSep'], out=ws['sep'][:b], pairmin=ws['pairmin'][:b])
            eng.speed(cpts, tf, E, -1.0, max_speed2, nveh=nv, out=ws['maxspeed'][:b])
            ready = torch.cuda.Event()
            ready.record(main)
            prev_ready = ready
            # the two result blocks leave on two copy streams (two copy engines share the link)
            with torch.cuda.stream(side2):
                side2.wait_event(ready)
                out['maxspeed'][lo:hi].copy_(ws['maxspeed'][:b].view(b, -1), non_blocking=True)
                done2 = torch.cuda.Event()
                done2.record(side2)
            with torch.cuda.stream(side):
                side.wait_event(ready)
                out['pairmin'][lo:hi].copy_(ws['pairmin'][:b], non_blocking=True)
                side.wait_event(done2)
                ws['done'] = torch.cuda.Event()
                ws['done'].record(side)
        side.synchronize()
        main.synchronize()
        self.workspace = {'key': None, 'sep': sw['sets'][(k & 1)]['sep']}
        return {'pairmin': out['pairmin'].numpy(), 'maxspeed': out['maxspeed'].numpy()}

    # -- cost callables (A14, optimization.py:287-308) -----------------------
    def _objective(self, x, kind):
        eng = self._engine(with_obstacles=False)
        E = _deg_elev()
        d_x = eng.upload(x)
        cpts, tf = eng.assemble(d_x, E)
        out = torch.empty((d_x.shape[0],), dtype=torch.float64, device=eng.device)
        plan = eng.plan(E)
        if kind == 'euclidean':
            _engine._capi.call("bez_objective_euclidean", plan.handle, _engine._ptr(cpts), int(d_x.shape[0]),
                               eng.N, eng.numVeh, _engine._ptr(out), _engine._stream())
        else:
            # the reference passes model['tf'] here even for time-optimal problems
            tfm = torch.full_like(tf, float(self.model['tf']))
            _engine._capi.call("bez_objective_accel", plan.handle, _engine._ptr(cpts), _engine._ptr(tfm),
                               int(d_x.shape[0]), eng.N, eng.numVeh, _engine._ptr(out), _engine._stream())
        return float(out[0].item())

    def euclideanObjective(self, x):
        """optimization.py:287-292 -> _euclideanObjective (:462-489)"""
        return self._objective(x, 'euclidean')

    def accelObjective(self, x):
        """optimization.py:294-300 -> _minAccelObjective (:503-519)"""
        return self._objective(x, 'accel')

    def jerkObjective(self, x):
        """optimization.py:302-308: the reference's _minJerkObjective is an @njit
        over Python objects and raises a numba TypingError when called."""
        raise NotImplementedError("the reference's jerk objective cannot run (numba TypingError); not provided")
