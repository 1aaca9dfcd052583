"""End-to-end drop-in check: SciPy's SLSQP driven by the GPU callables on
Example1 (Examples/Example1_DubinsCarTimeOptimal.py:95-148, BASELINE config C2).
SURVEY section 4 golden: tf = 2.4276431891903045 (elev 0), nit 22."""
import numpy as np
import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]


def _problem(gopt):
    return gopt.BezOptimization(numVeh=2, dimension=2, degree=10, minimizeGoal='TimeOpt', maxSep=1,
                                maxSpeed=5, maxAngRate=1, initPoints=[(0, 5), (3, 0)],
                                finalPoints=[(8, 4), (7, 10)], initSpeeds=[1, 1], finalSpeeds=[1, 1],
                                initAngs=[0, np.pi / 2], finalAngs=[0, np.pi / 2],
                                pointObstacles=[[3, 2], [6, 7]])


def _own_sep(bezopt, elev):
    """the example's own separation closure: vehicles only, its own elevation
    (Example1_DubinsCarTimeOptimal.py:19-49,125-126)"""
    from optimalbeziertrajectorygeneration_b200 import optimization as gopt
    veh_only = gopt.BezOptimization(numVeh=2, dimension=2, degree=10, minimizeGoal='TimeOpt', maxSep=1,
                                    initPoints=[(0, 5), (3, 0)], finalPoints=[(8, 4), (7, 10)],
                                    initSpeeds=[1, 1], finalSpeeds=[1, 1], initAngs=[0, np.pi / 2],
                                    finalAngs=[0, np.pi / 2])

    def fun(x):
        old = gopt.DEG_ELEV
        gopt.DEG_ELEV = elev
        try:
            return veh_only.temporalSeparationConstraints(x)
        finally:
            gopt.DEG_ELEV = old

    def jac(x):
        old = gopt.DEG_ELEV
        gopt.DEG_ELEV = elev
        try:
            return veh_only.temporalSeparationConstraints_jac(x)
        finally:
            gopt.DEG_ELEV = old
    return fun, jac


@pytest.mark.parametrize("with_jac", [False, True])
def test_example1_slsqp(with_jac):
    import scipy.optimize as sop
    from optimalbeziertrajectorygeneration_b200 import optimization as gopt
    gopt.DEG_ELEV = 0
    bezopt = _problem(gopt)
    x0 = bezopt.generateGuess(std=0)
    sep, sep_jac = _own_sep(bezopt, 0)
    cons = [{'type': 'ineq', 'fun': sep},
            {'type': 'ineq', 'fun': bezopt.maxSpeedConstraints},
            {'type': 'ineq', 'fun': bezopt.maxAngularRateConstraints},
            {'type': 'ineq', 'fun': lambda x: x[-1]}]
    if with_jac:
        cons[0]['jac'] = sep_jac
        cons[1]['jac'] = bezopt.maxSpeedConstraints_jac
        cons[2]['jac'] = bezopt.maxAngularRateConstraints_jac
    res = sop.minimize(bezopt.objectiveFunction, x0=x0, method='SLSQP', constraints=cons,
                       jac=bezopt.objectiveFunction_jac if with_jac else None,
                       options={'maxiter': 250, 'disp': False})
    assert res.success
    assert res.fun == pytest.approx(2.4276431891903045, rel=2e-5)
    # feasible at the solution
    assert sep(res.x).min() > -1e-6 and bezopt.maxSpeedConstraints(res.x).min() > -1e-6
    assert bezopt.maxAngularRateConstraints(res.x).min() > -1e-6
