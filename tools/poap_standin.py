"""Minimal stand-in for the POAP event loop (test / golden-generation infrastructure only).

POAP is an external dependency of pySOT (`setup.py:16`) that is neither vendored in /root/reference nor
installed in this image, and there is no network.  pySOT's strategies only rely on the small protocol described
in `docs/poap.rst:34-101` and visible at their call sites (`surrogate_strategy.py:249-439`,
`controller/controller.py:42-123`); this module honours exactly that protocol:

  Proposal(action, *args)       .accepted .args .record .add_callback .accept() .reject()
  EvalRecord(params)            .params .value .status .is_done .is_completed ... .add_callback, free attributes
  BaseStrategy                  propose_eval / propose_terminate helpers
  SerialController(objective)   run(): propose -> evaluate inline -> callbacks
  SimAsyncController            deterministic emulation of ThreadController with W workers: proposals are
                                accepted while a worker is free, otherwise the OLDEST running evaluation
                                completes (so up to W-1 points are pending when a proposal is generated)
  ThreadController              real threads: workers evaluate concurrently and report through a message queue that
  BasicWorkerThread             only the controller thread drains, so strategy / surrogate code runs on ONE thread
                                while objectives run on the others (docs/poap.rst:54-63, tests/test_strategies.py:60-82)

Install with `install()` BEFORE importing pySOT so that `from poap.strategy import BaseStrategy, Proposal`
resolves here.
"""
import queue
import sys
import threading
import types


class Proposal:
    def __init__(self, action, *args):
        self.action, self.args = action, args
        self.accepted = False
        self.callbacks = []
        self.record = None

    def add_callback(self, cb):
        self.callbacks.append(cb)

    def copy(self):
        p = Proposal(self.action, *self.args)
        p.callbacks = list(self.callbacks)
        return p

    def accept(self):
        self.accepted = True
        for cb in self.callbacks:
            cb(self)

    def reject(self):
        self.accepted = False
        for cb in self.callbacks:
            cb(self)


class EvalRecord:
    def __init__(self, params, status="pending"):
        self.params, self.status, self.value = params, status, None
        self.callbacks = []

    def add_callback(self, cb):
        self.callbacks.append(cb)

    def update(self):
        for cb in self.callbacks:
            cb(self)

    def running(self):
        self.status = "running"
        self.update()

    def complete(self, value):
        self.status, self.value = "completed", value
        self.update()

    def kill(self):
        self.status = "killed"
        self.update()

    def cancel(self):
        self.status = "cancelled"
        self.update()

    is_pending = property(lambda s: s.status == "pending")
    is_running = property(lambda s: s.status == "running")
    is_completed = property(lambda s: s.status == "completed")
    is_killed = property(lambda s: s.status == "killed")
    is_cancelled = property(lambda s: s.status == "cancelled")
    is_done = property(lambda s: s.status in ("completed", "killed", "cancelled"))


class BaseStrategy:
    def propose_eval(self, *args):
        return Proposal("eval", *args)

    def propose_terminate(self):
        return Proposal("terminate")


class RetryStrategy:  # only named at import time by pySOT/strategy/random_strategy.py
    pass


class Controller:
    def __init__(self):
        self.strategy = None
        self.fevals = []
        self.feval_callbacks = []
        self.term_callbacks = []

    def add_feval_callback(self, cb):
        self.feval_callbacks.append(cb)

    def add_term_callback(self, cb):
        self.term_callbacks.append(cb)

    def new_feval(self, params):
        r = EvalRecord(params)
        self.fevals.append(r)
        for cb in self.feval_callbacks:
            cb(r)
        return r

    def best_point(self):
        done = [r for r in self.fevals if r.is_completed]
        return min(done, key=lambda r: r.value) if done else None

    def _terminate(self):
        for cb in self.term_callbacks:
            cb()
        return self.best_point()


class SerialController(Controller):
    def __init__(self, objective):
        super().__init__()
        self.objective = objective

    def run(self):
        while True:
            p = self.strategy.propose_action()
            if p is None:
                raise RuntimeError("serial controller: strategy proposed nothing")
            if p.action == "terminate":
                p.accept()
                return self._terminate()
            r = self.new_feval(p.args)
            p.record = r
            p.accept()
            r.running()
            r.complete(self.objective(*p.args))


class SimAsyncController(Controller):
    """Deterministic ThreadController emulation: `workers` evaluations in flight, oldest completes first."""

    def __init__(self, objective, workers):
        super().__init__()
        self.objective, self.workers = objective, workers
        self.running = []

    def _complete_oldest(self):
        r = self.running.pop(0)
        r.complete(self.objective(*r.params))

    def run(self):
        while True:
            # Ask for a proposal only while a worker is free.  pySOT's asynchronous strategies DROP a rejected
            # adaptive proposal (surrogate_strategy.py:399-406) and SOPStrategy then leaks the center it was
            # generated from (sop_strategy.py:257-262) until generate_evals raises IndexError, so a controller that
            # proposed with all workers busy would break the reference itself.
            if len(self.running) >= self.workers:
                self._complete_oldest()
                continue
            p = self.strategy.propose_action()
            if p is None:
                if not self.running:
                    raise RuntimeError("async controller: nothing proposed and nothing running")
                self._complete_oldest()
                continue
            if p.action == "terminate":
                p.accept()
                return self._terminate()
            if len(self.running) >= self.workers:
                p.reject()
                self._complete_oldest()
                continue
            r = self.new_feval(p.args)
            p.record = r
            p.accept()
            r.running()
            self.running.append(r)


class ThreadController(Controller):
    """Threaded controller (docs/poap.rst:54-63).  Idle workers wait in ``self.workers``; anything that touches a record or
    the strategy is posted as a closure to ``self.messages`` and executed by ``run()`` on the caller's thread.

    Loop: drain the queued messages, ask the strategy for a proposal;  no proposal -> block for one message;
    ``terminate`` -> accept and stop;  ``eval`` with an idle worker -> new record, accept, hand the record to the worker;
    anything else -> reject and block for one message (so a strategy that always proposes does not spin)."""

    def __init__(self):
        super().__init__()
        self.workers = queue.Queue()
        self.messages = queue.Queue()
        self._threads = []

    def add_worker(self, worker):
        self.workers.put(worker)

    def launch_worker(self, worker, daemon=True):
        worker.daemon = daemon
        self._threads.append(worker)
        self.add_worker(worker)
        worker.start()

    def add_message(self, message=None):
        self.messages.put(message)

    def can_work(self):
        return not self.workers.empty()

    def _run_message(self):
        message = self.messages.get()  # blocks
        if message is not None:
            message()

    def _run_queued_messages(self):
        while True:
            try:
                message = self.messages.get_nowait()
            except queue.Empty:
                return
            if message is not None:
                message()

    def run(self):
        try:
            while True:
                self._run_queued_messages()
                p = self.strategy.propose_action()
                if p is None:
                    self._run_message()
                elif p.action == "terminate":
                    p.accept()
                    return self._terminate()
                elif p.action == "eval" and self.can_work():
                    worker = self.workers.get_nowait()
                    r = self.new_feval(p.args)
                    r.worker = worker
                    p.record = r
                    p.accept()
                    worker.eval(r)
                else:
                    p.reject()
                    self._run_message()
        finally:
            for w in self._threads:
                w.terminate()


class BasicWorkerThread(threading.Thread):
    """Worker that calls a Python objective on its own thread (tests/test_strategies.py:62-64).  Record updates are
    never made here: they are posted to the controller as messages.  An objective that raises kills the record."""

    def __init__(self, controller, objective):
        super().__init__()
        self.controller, self.objective = controller, objective
        self.requests = queue.Queue()

    def eval(self, record):
        self.requests.put(("eval", record))

    def kill(self, record):  # a Python call cannot be interrupted; POAP's basic worker ignores kill requests too
        pass

    def terminate(self):
        self.requests.put(("terminate", None))

    def run(self):
        while True:
            what, record = self.requests.get()
            if what == "terminate":
                return
            self.controller.add_message(record.running)
            try:
                value = self.objective(*record.params)
            except Exception:
                self.controller.add_message(record.kill)
            else:
                self.controller.add_message(lambda r=record, v=value: r.complete(v))
            self.controller.add_worker(self)  # idle again, after its result was posted


def install():
    """Register this module as `poap`, `poap.strategy`, `poap.controller`."""
    me = sys.modules[__name__]
    poap = types.ModuleType("poap")
    poap.strategy, poap.controller = me, me
    sys.modules.update({"poap": poap, "poap.strategy": me, "poap.controller": me})
    return poap
