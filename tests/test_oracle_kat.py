"""Pin the restated dopri5 + RHS oracle on the losses the reference logged itself.

Default run: the four AP-2Hz rows plus two step-protocol rows (about a minute).  The complete
92-row sweep runs with ``IKR_FULL_KAT=1`` (about 10 minutes, 1 thread)."""
import os

import pytest
import torch

from tests import kat

TOL = 5e-5     # fp32-state noise floor: tolerances sit at fp32 eps, ~25 % of steps are rejected

FAST = [(s, 0) for s in kat.STUDIES] + [('s1', 4), ('d2', 11)]


def _rows(study):
    return [r for r in kat.KAT[study] if r['inputs_present']]


@pytest.mark.parametrize('study,idx', FAST)
def test_logged_loss_fast(study, idx):
    torch.set_num_threads(1)
    row = kat.KAT[study][idx]
    assert row['inputs_present']
    loss, _, _, _ = kat.oracle_case(study, row)
    assert abs(loss - row['loss']) < TOL, (study, row, loss)


@pytest.mark.slow
@pytest.mark.skipif(not os.environ.get('IKR_FULL_KAT'), reason='set IKR_FULL_KAT=1')
@pytest.mark.parametrize('study', kat.STUDIES)
def test_logged_loss_full(study):
    torch.set_num_threads(1)
    worst = 0.0
    for row in _rows(study):
        loss, _, _, _ = kat.oracle_case(study, row)
        worst = max(worst, abs(loss - row['loss']))
        assert abs(loss - row['loss']) < TOL, (study, row, loss)
    print(study, 'worst |delta|', worst)
