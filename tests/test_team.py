"""The launcher threads of a multi-device context (lambda_blas_b200/csrc/lbk_team.hpp), on the CPU: every job runs exactly
once on every rank, statuses come back in rank order, sleeping members are woken."""
import ctypes as C

import pytest


def run(blas, nranks, jobs, fail_at=-1, pause_us=0):
    from lambda_blas_b200 import _ffi
    total, first = C.c_int64(), C.c_int()
    _ffi.check(_ffi.lib().lbk_team_selftest(nranks, jobs, fail_at, pause_us, C.byref(total), C.byref(first)))
    return total.value, first.value


@pytest.mark.parametrize("nranks", [1, 2, 3, 8])
def test_every_rank_runs_every_job_once(blas, nranks):
    jobs = 2000
    total, first = run(blas, nranks, jobs)
    assert total == (jobs * (jobs + 1) // 2) * (nranks * (nranks + 1) // 2) and first == 0


def test_members_that_went_to_sleep_are_woken(blas):
    total, first = run(blas, 4, 12, pause_us=3000)          # 3 ms between jobs: members spin 200 us, then sleep
    assert total == (12 * 13 // 2) * 10 and first == 0


def test_first_failing_rank_is_reported(blas):
    total, first = run(blas, 5, 50, fail_at=7)
    assert first == 100                                    # rank order: rank 0's status
    assert total == (50 * 51 // 2) * 15                    # a failing job does not stop the team


def test_empty_team(blas):
    assert run(blas, 3, 0) == (0, 0)
