// lbk_pdl.hpp -- which kernels of a stream may overlap (programmatic dependent launch), decided on the host.
//
// Pure host logic, no CUDA types: lbk_lib.cu asks it how to launch each streaming kernel (unit-stride, and since round 2 the
// general-stride ones, whose operands are stated as whole spans), and the CPU
// tests drive it directly through the lbk_pdl_sim_* entry points (tests/test_pdl_plan.py).
//
// Every unit-stride kernel releases its dependents as its first instruction
// (griddepcontrol.launch_dependents), so the NEXT unit-stride kernel, when it is launched with the
// programmatic-serialization attribute, fills SMs while this one's last blocks retire.  What may still be
// running when a kernel starts is the GROUP: the chain of PDL launches since the last ordinary launch (or
// any other stream operation), with everything its members read and write.  A new kernel is planned
// against the group:
//   * it reads something the group writes          -> true dependency: PDL all the same, but the kernel
//     executes griddepcontrol.wait before its first load and only then releases its own dependents --
//     the semantics of an ordinary launch with the launch latency hidden (its blocks are resident and
//     past their prologue when the predecessors finish);
//   * it writes something the group reads or writes -> PDL, but it executes griddepcontrol.wait (all
//     predecessors complete and visible) before its first store; its loads still overlap;
//   * otherwise                                     -> PDL, griddepcontrol.wait only before it exits
//     (completion stays in stream order).
// A reduction touches the shared scratch, its result and the mailbox only after the wait, so only what it
// READS can collide with the group.
//
// A reduction can also release its dependents LATE: after its last load and after its own wait.  The group
// then shrinks to the reduction's result, so the next kernel may overwrite the reduction's operands, or read
// what an earlier member of the chain wrote, and still run under the reduction's tail.  Which is better
// depends on the NEXT call, which is unknown at launch time, so it is predicted from what followed the same
// call (routine, operands) the last time; a wrong guess costs the overlap, never correctness.
#pragma once

#include <stdint.h>

namespace lbk_pdl {

enum Mode { NONE = 0, WAIT_AT_EXIT = 1, WAIT_BEFORE_STORE = 2, WAIT_BEFORE_LOAD = 3 };

struct Range { const char *lo, *hi; };

struct Group {
    static constexpr int kCap = 24;
    bool open = false;
    int nr = 0, nw = 0;
    Range r[kCap], w[kCap];
};

// One unit-stride kernel about to be launched: what it reads and writes (pointer, bytes).
struct Op {
    bool reduction = false;          // writes only after griddepcontrol.wait
    int rid = 0;                     // routine (predictor key, reductions only)
    const void *rd[2] = {nullptr, nullptr};
    int64_t rdn[2] = {0, 0};
    int nrd = 0;
    const void *wr[2] = {nullptr, nullptr};
    int64_t wrn[2] = {0, 0};
    int nwr = 0;
};

struct Plan {
    Mode mode = NONE;
    bool late = false;               // reductions: release dependents after the loads and the wait
    int slot = 0;                    // predictor slot of this reduction
};

struct State {
    bool enabled = true;
    bool hold_stores = true;         // false: an anti/output dependency is treated like a true one (experiments)
    bool hold_loads = true;          // false: a true dependency gets an ordinary launch (experiments)
    Group grp;
    // the predictor: did the kernel that followed this call last time collide with it or its chain?
    static constexpr int kSlots = 64;
    struct Slot { const void *x = nullptr, *y = nullptr; int rid = -1; bool late = false; } slot[kSlots];
    // the reduction launched last, until the next unit-stride kernel tells what it was followed by;
    // `hyp` is the group as it would stand had that reduction released its dependents at once
    bool pending = false;
    int pending_slot = 0;
    Group hyp;
};

inline bool hits(const Range *set, int n, const void *p, int64_t nbytes)
{
    if (!p || nbytes <= 0) return false;
    const char *lo = (const char *)p, *hi = lo + nbytes;
    for (int i = 0; i < n; ++i)
        if (lo < set[i].hi && set[i].lo < hi) return true;
    return false;
}

inline Mode decide(const Group &g, const Op &op)
{
    if (!g.open) return NONE;
    for (int i = 0; i < op.nrd; ++i)
        if (hits(g.w, g.nw, op.rd[i], op.rdn[i])) return WAIT_BEFORE_LOAD;   // reads what a predecessor writes
    if (op.reduction) return WAIT_AT_EXIT;
    for (int i = 0; i < op.nwr; ++i)
        if (hits(g.r, g.nr, op.wr[i], op.wrn[i]) || hits(g.w, g.nw, op.wr[i], op.wrn[i])) return WAIT_BEFORE_STORE;
    return WAIT_AT_EXIT;
}

inline void add(Range *set, int *n, const void *p, int64_t nbytes, bool *full)
{
    if (!p || nbytes <= 0) return;
    const char *lo = (const char *)p, *hi = lo + nbytes;
    for (int i = 0; i < *n; ++i)
        if (set[i].lo <= lo && hi <= set[i].hi) return;                      // already covered
    if (*n == Group::kCap) { *full = true; return; }
    set[(*n)++] = Range{lo, hi};
}

inline int slot_of(int rid, const void *x, const void *y)
{
    return (int)((((uintptr_t)x >> 8) * 31u + ((uintptr_t)y >> 8) * 17u + (unsigned)rid) % State::kSlots);
}

// Any other operation on the stream (copy, event, NCCL, graph launch): the next kernel is launched the ordinary way.
inline void interrupt(State &s) { s.grp.open = false; }

// How to launch `op` behind everything enqueued so far.  Also teaches the predictor what the last
// reduction was followed by.
inline Plan plan(State &s, const Op &op)
{
    Plan p;
    if (s.pending) {      // released at once, would the last reduction have held this kernel back?
        s.slot[s.pending_slot].late = decide(s.hyp, op) != WAIT_AT_EXIT;
        s.pending = false;
    }
    p.mode = s.enabled ? decide(s.grp, op) : NONE;
    if (p.mode == WAIT_BEFORE_STORE && !s.hold_stores) p.mode = WAIT_BEFORE_LOAD;
    if (p.mode == WAIT_BEFORE_LOAD && !s.hold_loads) p.mode = NONE;
    if (op.reduction) {
        p.slot = slot_of(op.rid, op.rd[0], op.rd[1]);
        State::Slot &sl = s.slot[p.slot];
        if (sl.x != op.rd[0] || sl.y != op.rd[1] || sl.rid != op.rid) { sl.x = op.rd[0]; sl.y = op.rd[1]; sl.rid = op.rid; sl.late = false; }
        p.late = sl.late;
    }
    return p;
}

// The kernel has been launched as planned: it joins (or restarts) the group.
inline void commit(State &s, const Op &op, const Plan &p)
{
    Group &g = s.grp;
    // an ordinary launch, or one that waits before its first load and only then releases its dependents,
    // starts (in effect) after everything before it
    if (p.mode == NONE || p.mode == WAIT_BEFORE_LOAD) { g.nr = g.nw = 0; }
    bool full = false;
    if (op.reduction) {                                // the group had this reduction released its dependents at once
        Group &h = s.hyp;
        h = g;
        bool hfull = false;
        for (int i = 0; i < op.nrd; ++i) add(h.r, &h.nr, op.rd[i], op.rdn[i], &hfull);
        for (int i = 0; i < op.nwr; ++i) add(h.w, &h.nw, op.wr[i], op.wrn[i], &hfull);
        h.open = !hfull;
        s.pending = true;
        s.pending_slot = p.slot;
    }
    if (p.late) { g.nr = g.nw = 0; }                   // its predecessors are complete and its loads done: only its result is in flight
    for (int i = 0; i < op.nrd && !p.late; ++i) add(g.r, &g.nr, op.rd[i], op.rdn[i], &full);
    for (int i = 0; i < op.nwr; ++i) add(g.w, &g.nw, op.wr[i], op.wrn[i], &full);
    g.open = !full;                                    // a full table ends the chain: the next launch is ordinary
}

}  // namespace lbk_pdl
