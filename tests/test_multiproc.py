"""Multi-process semantics over Gloo, world_size 2 and 4 on one machine (SURVEY.md §4 tier 2)."""
import pytest

from mp_util import run_workers


@pytest.mark.parametrize("world", [2, 4])
def test_broadcast_parameters(world):
    assert all(run_workers(world, "mp_cases", "broadcast_parameters"))


@pytest.mark.parametrize("world,opt,passes", [(2, "sgd", 1), (4, "adam", 1), (2, "adam", 2)])
def test_dp_step_equals_single_process(world, opt, passes):
    assert all(run_workers(world, "mp_cases", "dp_equals_single", (opt, passes)))


def test_optimizer_edge_cases():
    assert all(run_workers(2, "mp_cases", "optimizer_edge_cases"))


@pytest.mark.parametrize("world", [2, 3])
def test_collectives(world):
    assert all(run_workers(world, "mp_cases", "collectives"))


def test_broadcast_optimizer_state():
    assert all(run_workers(2, "mp_cases", "broadcast_optimizer_state"))


def test_sync_batch_norm():
    assert all(run_workers(2, "mp_cases", "sync_batch_norm"))


def test_timeline_and_elastic(tmp_path):
    assert all(run_workers(2, "mp_cases", "timeline_and_elastic", (str(tmp_path),)))


def test_optimizer_options():
    assert all(run_workers(2, "mp_cases", "optimizer_options"))
