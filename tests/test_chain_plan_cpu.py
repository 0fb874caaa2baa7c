"""CPU: the row-block schedule of the fused DiT chain (dit_chain.cu, dit_chain_plan) through its host-only C entry point.
The schedule decides how many rounds a launch takes; before it became two-phase, a block count one above a multiple of the resident
CTA pairs cost a whole extra round -- and BASELINE.json configs[3] sits exactly there (593 = 8 x 74 + 1 row blocks on one GPU,
297 / 149 / 75 per rank on 2 / 4 / 8 GPUs)."""
import ctypes

import pytest

import b200tts  # noqa: F401
from b200tts import capi

PAIRS = 74                      # 148 SMs
COST = {1: 1.0, 2: 0.55, 4: 0.30, 8: 0.17}


def plan(nrb, pairs=PAIRS):
    out = (ctypes.c_int * 5)()
    assert capi.load_library().b200tts_debug_chain_plan(nrb, pairs, out) == 0
    team, teams, nrb0, rem, team1 = list(out)
    return team, teams, nrb0, rem, team1


def modelled(p):
    team, teams, nrb0, rem, team1 = p
    rounds0 = -(-nrb0 // teams)
    return rounds0 * COST[team] + (COST[team1] if rem else 0.0)


@pytest.mark.parametrize("pairs", [74, 66, 9, 8])
def test_every_row_block_is_scheduled_exactly_once_and_teams_fit(pairs):
    for nrb in list(range(1, 700)) + [1186, 4737]:
        team, teams, nrb0, rem, team1 = plan(nrb, pairs)
        assert team in (1, 2, 4, 8) and team1 in (1, 2, 4, 8)
        assert 1 <= teams and teams * team <= pairs                   # phase 0 fits the resident pairs
        assert nrb0 + rem == nrb and nrb0 >= 1 and rem >= 0
        if rem:
            assert nrb0 % teams == 0                                   # phase 0 = whole rounds only
            assert rem * team1 <= pairs and team1 >= team              # phase 1: one team per remaining block
            assert rem < teams                                         # ... and it really is a remainder
        covered = set()
        for t in range(teams):                                         # the kernel's loops (get_phase): rb = t, t + teams, ... < nrb0
            covered.update(range(t, nrb0, teams))
        covered.update(nrb0 + t for t in range(rem))
        assert covered == set(range(nrb))


def test_known_plans_of_the_benchmark_shapes():
    assert plan(9) == (8, 9, 9, 0, 1)                   # one config-3 utterance: nine blocks, a team of 8 each
    assert plan(71) == (1, 71, 71, 0, 1)                # eight uniform utterances: one round, a pair per block
    assert plan(593) == (1, 74, 592, 1, 8)              # configs[3] on one GPU: 8 rounds + one block shared by 8 pairs
    assert plan(297) == (1, 74, 296, 1, 8)              # ... per rank on 2 GPUs
    assert plan(149) == (1, 74, 148, 1, 8)              # ... 4 GPUs
    assert plan(75) == (1, 74, 74, 1, 8)                # ... 8 GPUs
    assert plan(40) == (2, 37, 37, 3, 8)                # half-empty: teams of 2 fill the chip, the rest in teams of 8


def test_no_cliff_one_block_above_a_multiple_of_the_pairs():
    for k in range(1, 12):
        at, above = modelled(plan(k * PAIRS)), modelled(plan(k * PAIRS + 1))
        assert at == pytest.approx(k * 1.0)
        assert above <= k + 0.17 + 1e-9                 # the odd block costs a T = 8 round, not a whole one
    # and the modelled time never beats the work bound (nrb / pairs rounds at T = 1 efficiency) once the chip is full
    for nrb in range(PAIRS, 700):
        assert modelled(plan(nrb)) >= nrb / PAIRS * 0.999


def test_bad_arguments_are_refused():
    out = (ctypes.c_int * 5)()
    lib = capi.load_library()
    assert lib.b200tts_debug_chain_plan(0, 74, out) != 0
    assert lib.b200tts_debug_chain_plan(10, 0, out) != 0
    assert lib.b200tts_debug_chain_plan(10, 74, None) != 0
