"""Barrier-protocol model of the attention-forward kernels (tools/protocol_sim.py): the shipped
protocols must survive randomised latencies without stale reads, overwritten live buffers,
mbarrier parity aliasing or deadlock - and the model must notice when a protocol is broken."""
import sys
from pathlib import Path

import pytest

sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tools"))
import protocol_sim as P  # noqa: E402


@pytest.mark.parametrize("mode", ["pp", "pp16", "wg1", "wg2"])
def test_protocol_is_hazard_free(mode):
    assert P.run(mode, trials=40)


@pytest.mark.parametrize("mode,mutate", [("wg1", "no_e_bar"), ("wg2", "no_e_bar"), ("pp", "no_q_wait"),
                                         ("wg2", "no_q_wait"), ("wg1", "ring4")])
def test_model_detects_broken_protocols(mode, mutate):
    with pytest.raises(P.Hazard):
        P.run(mode, trials=60, mutate=mutate)


@pytest.mark.parametrize("mutate", ["no_p_free", "no_s_free"])
def test_redundant_waits_are_implied_by_the_in_order_tensor_pipe(mutate):
    # S(g) is queued behind PV(g-2) and issue_pv(g-2) already waited for block g-2's softmax, so
    # these two waits in the kernels are belt and braces; the model agrees.
    assert P.run("pp", trials=30, mutate=mutate)
    assert P.run("wg2", trials=30, mutate=mutate)


@pytest.mark.parametrize("mode", ["bwd", "bwd_split"])
def test_backward_protocols_are_hazard_free(mode):
    assert P.run_bwd(mode, trials=30)


@pytest.mark.parametrize("mutate", ["no_hfree", "no_g_wait"])
def test_model_detects_broken_backward_split(mutate):
    with pytest.raises(P.Hazard):
        P.run_bwd("bwd_split", trials=60, mutate=mutate)
