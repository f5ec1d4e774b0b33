"""The launch planner (lambda_blas_b200/csrc/lbk_pdl.hpp) on the CPU: which consecutive kernels may overlap.

Driven through lbk_pdl_sim_* (include/lbk.h): no device involved.  mode 0 = ordinary launch, 1 = programmatic
dependent launch with the wait for the predecessors at exit, 2 = before the first store, 3 = before the first load
(a true dependency: ordinary-launch semantics with the launch latency hidden)."""
import ctypes as C

import pytest

X, Y, Z, W, OUT = 0x10000000, 0x20000000, 0x30000000, 0x40000000, 0x7f000000
N = 1 << 20                      # bytes per vector
DOT, NRM2, ASUM = 0, 3, 2        # routine ids (only the predictor's key)


class Sim:
    def __init__(self, blas):
        from lambda_blas_b200 import _ffi
        self.L = _ffi.lib()
        self.h = C.c_void_p()
        assert self.L.lbk_pdl_sim_create(C.byref(self.h)) == 0

    def close(self):
        self.L.lbk_pdl_sim_destroy(self.h)

    def op(self, reads, writes, reduction=False, rid=0):
        rd = (C.c_int64 * 4)(*[v for a in reads for v in (a if isinstance(a, tuple) else (a, N))] + [0] * (4 - 2 * len(reads)))
        wr = (C.c_int64 * 4)(*[v for a in writes for v in (a if isinstance(a, tuple) else (a, N))] + [0] * (4 - 2 * len(writes)))
        mode, late = C.c_int(), C.c_int()
        assert self.L.lbk_pdl_sim_op(self.h, int(reduction), rid, rd, len(reads), wr, len(writes), C.byref(mode), C.byref(late)) == 0
        return mode.value, late.value

    def red(self, rid, *reads):
        return self.op(list(reads), [(OUT, 8)], True, rid)

    def interrupt(self):
        self.L.lbk_pdl_sim_interrupt(self.h)


@pytest.fixture
def sim(blas):
    s = Sim(blas)
    yield s
    s.close()


def test_first_kernel_and_interrupts_are_ordinary(sim):
    assert sim.op([X], [Y])[0] == 0                 # nothing in flight to overlap
    assert sim.op([Z], [W])[0] == 1                 # independent: overlaps
    sim.interrupt()                                 # a copy / event / allocation on the stream
    assert sim.op([X], [Z])[0] == 0


def test_hazard_classes(sim):
    sim.op([X], [Y])                                # y = f(x)
    assert sim.op([Y], [Z])[0] == 3                 # read after write: waits before its first load
    assert sim.op([X], [W])[0] == 1                 # reads x (only read so far), writes w: free
    assert sim.op([(X + 6 * N, N)], [X])[0] == 2    # write after read (x): hold the stores
    assert sim.op([(X + 6 * N, N)], [(X + N // 2, N)])[0] == 2   # write after write, partial overlap
    assert sim.op([Z], [(Y + 6 * N, N)])[0] == 3    # z was written by a kernel of the chain
    assert sim.op([(X + 4 * N, N)], [(Y + 4 * N, N)])[0] == 1    # disjoint windows


def test_chain_is_checked_as_a_whole(sim):
    sim.op([X], [Y])                                # ordinary
    assert sim.op([Z], [W])[0] == 1                 # overlaps the first
    # the first kernel may still be running when a third starts: its write of y still counts
    assert sim.op([Y], [(X + 8 * N, N)])[0] == 3
    # ... and a kernel that waited before its first load restarts the chain: nothing older is in flight behind it
    assert sim.op([W], [X])[0] == 1


def test_reduction_reads_only(sim):
    sim.op([X], [Y])
    assert sim.red(NRM2, X)[0] == 1                 # reads what the map only read
    assert sim.red(NRM2, Y)[0] == 3                 # reads what the map wrote
    # a reduction writes its result after the wait: two reductions into the same slot still overlap
    assert sim.red(ASUM, X)[0] == 1


def test_predictor_learns_the_bench_step(sim):
    """sdot(x,y) -> saxpy (writes y) -> snrm2(x), repeated: from the second step on every boundary overlaps."""
    history = []
    for step in range(4):
        history.append((sim.red(DOT, X, Y), sim.op([X, Y], [Y]), sim.red(NRM2, X)))
    (d0, a0, n0), (d1, a1, n1) = history[0], history[1]
    assert d0 == (0, 0) and a0[0] == 2 and n0 == (1, 0)       # first pass: released at once, saxpy holds its stores
    # second pass: sdot reads y that saxpy wrote and snrm2 was released at once -> it waits before its first
    # load, but both reductions have learnt to release late
    assert d1 == (3, 1) and a1[0] == 1 and n1 == (1, 1)
    for d, a, n in history[2:]:
        assert d == (1, 1) and a[0] == 1 and n == (1, 1)       # steady state: all three overlap, no held stores


def test_predictor_keeps_independent_reductions_early(sim):
    for _ in range(3):
        assert sim.red(NRM2, X)[1] == 0 and sim.red(ASUM, X)[1] == 0 and sim.red(DOT, X, Y)[1] == 0


def test_full_table_ends_the_chain(sim):
    sim.op([X], [Y])
    modes = [sim.op([(Z + 2 * k * N, N)], [(W + 2 * k * N, N)])[0] for k in range(40)]
    assert modes[0] == 1 and 0 in modes             # the chain is cut when the table is full, then starts again
    assert modes[modes.index(0) + 1] == 1


def test_disabled(sim):
    sim.L.lbk_pdl_sim_enable(sim.h, 0)
    assert [sim.op([X], [Y])[0], sim.op([Z], [W])[0], sim.red(NRM2, X)[0], sim.op([Y], [Z])[0]] == [0, 0, 0, 0]


def _overlap(a, b):
    return a[0] < b[0] + b[1] and b[0] < a[0] + a[1]


def test_random_sequences_never_admit_a_hazard(blas):
    """Model check: replay random kernel sequences through the planner and, from the MODES it hands out alone, keep the
    set of kernels that may physically still be running when the next one starts.  Whatever the planner's own
    bookkeeping (capacity, predictor, late release), no kernel may then read what a running kernel writes, nor
    write -- before its wait -- what a running kernel reads or writes."""
    import numpy as np
    rng = np.random.default_rng(2024)
    bases = [0x1000000 * (k + 1) for k in range(6)]
    for trial in range(200):
        sim = Sim(blas)
        flight = []                               # (reads, writes) of kernels that may still be running
        for step in range(int(rng.integers(5, 80))):
            if rng.random() < 0.08:
                sim.interrupt(); flight = None    # any other stream operation: the next launch must be ordinary
                continue
            def rng_range():
                b = bases[int(rng.integers(len(bases)))]
                off = int(rng.integers(0, 4)) * (N // 4)
                return (b + off, int(rng.integers(1, 5)) * (N // 4))
            reduction = rng.random() < 0.4
            reads = [rng_range() for _ in range(int(rng.integers(1, 3)))]
            writes = [(OUT + 16 * int(rng.integers(4)), 8)] if reduction else [rng_range() for _ in range(int(rng.integers(1, 3)))]
            if not reduction and rng.random() < 0.5:
                reads[0] = writes[0]              # an in-place update
            mode, late = sim.op(reads, writes, reduction, int(rng.integers(3)))
            if flight is None:
                assert mode == 0                  # after an interruption nothing may overlap
                flight = []
            if mode in (0, 3):
                flight = []                       # starts (in effect) after everything before it
            for fr, fw in flight:
                assert not any(_overlap(r, w) for r in reads for w in fw), "reads what a running kernel writes"
                if mode == 1 and not reduction:
                    assert not any(_overlap(w, o) for w in writes for o in fr + fw), "stores over a running kernel's operands"
            if late:
                assert reduction
                flight = [([], writes)]           # released after its loads and its predecessors: only its result is pending
            else:
                flight.append((reads, writes))
        sim.close()
