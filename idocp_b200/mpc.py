"""One control tick of a BATCH of MPC loops around OCPSolver (the caller one step outside the hot path).

idocp has no MPC class of its own: its consumers (idocp-sim, README.md:118-119) call, every control period,
    [popFrontContactStatus() once the first switching time has passed]          ocp_solver.cpp:174-194
    [pushBackContactStatus(next phase, time) to keep the horizon covered]
    updateSolution(t, q, v) a fixed number of times (the iterate of the previous tick is the warm start: it stays in place)
    u = getSolution(0).u,  K = getStateFeedbackGain(0)                           riccati_recursion_solver.cpp:254-260
`BatchedMPC.tick` is exactly that sequence for all instances at once; nothing numerical happens here."""
import numpy as np


class BatchedMPC:
    def __init__(self, solver, iterations=1, gait=None, line_search=False):
        """solver: idocp_b200.OCPSolver; gait(mpc, t): optional callback that pushes new contact phases
        (solver.pushBackContactStatus) so that the schedule keeps covering [t, t + T]."""
        self.solver, self.iterations, self.gait, self.line_search = solver, int(iterations), gait, bool(line_search)
        self.popped = 0

    def first_event_time(self):
        cs = self.solver.contact_sequence
        times = []
        if cs.numImpulseEvents() > 0:
            times.append(cs.impulseTime(0))
        if cs.numLiftEvents() > 0:
            times.append(cs.liftTime(0))
        return min(times) if times else None

    def advance_schedule(self, t):
        """Drop the phases whose switching time has passed (an event before t makes the discretisation ill-defined,
        ocp_discretizer.hxx:62-72) and let the gait callback extend the schedule."""
        while True:
            te = self.first_event_time()
            if te is None or te > t:
                break
            self.solver.popFrontContactStatus()
            self.popped += 1
        if self.gait is not None:
            self.gait(self, t)

    def tick(self, t, q, v, with_gain=False):
        """-> u0 (batch, 12) [, (Kq, Kv)]: the first control input of every instance after `iterations` Newton iterations
        from the previous tick's solution, and optionally the LQR feedback gain of the first stage."""
        self.advance_schedule(t)
        for _ in range(self.iterations):
            self.solver.updateSolution(t, q, v, self.line_search)
        u0 = np.asarray(self.solver.get(0, "u"), dtype=float).reshape(self.solver.B, -1)
        if with_gain:
            return u0, self.solver.getStateFeedbackGain(0)
        return u0
