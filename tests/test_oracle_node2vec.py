"""Pin the node2vec CPU oracle (oracle/n2v_oracle.py) against fixtures produced by the
real reference (tests/golden/make_golden_node2vec.py).  CPU only."""
import json
import os

import numpy as np
import pytest

from oracle import n2v_oracle as O
from conftest import GOLDEN, DATA

CASES = json.load(open(os.path.join(GOLDEN, "n2v_cases.json")))


def load_case(meta):
    z = np.load(os.path.join(GOLDEN, "n2v_%s.npz" % meta["name"]))
    g = O.load_graph(os.path.join(DATA, meta["file"]), meta["delimiter"], meta["weighted"],
                     meta["directed"])
    return z, g


def test_alias_setup_kats():
    for kat in json.load(open(os.path.join(GOLDEN, "alias_setup_kat.json"))):
        J, q = O.alias_setup(kat["probs"])
        assert [int(x) for x in J] == kat["J"]
        assert [float(x).hex() for x in q] == kat["q_hex"]
    # SURVEY §8(a2) literal KAT
    J, q = O.alias_setup([0.1, 0.2, 0.3, 0.4])
    assert list(J) == [3, 3, 0, 2] and list(q) == [0.4, 0.8, 1.0, 0.8000000000000003]


@pytest.mark.parametrize("meta", CASES, ids=[c["name"] for c in CASES])
def test_csr_matches_reference(meta):
    z, g = load_case(meta)
    for k in ("node_ids", "first_seen", "row_ptr", "col_idx"):
        assert np.array_equal(z[k], g[k]), k
    assert z["weights"].tobytes() == g["weights"].tobytes()


@pytest.mark.parametrize("meta", CASES, ids=[c["name"] for c in CASES])
def test_alias_tables_bit_exact(meta):
    z, g = load_case(meta)
    J, q = O.alias_nodes_flat(g)
    assert np.array_equal(J, z["an_J"])
    assert q.tobytes() == z["an_q"].tobytes()
    if meta["n_alias_edge_entries"] > 20000:
        pytest.skip("edge tables of this case are covered by the GPU parity test")
    off, J, q = O.alias_edges_flat(g, meta["p"], meta["q"])
    assert np.array_equal(off, z["ae_off"])
    assert np.array_equal(J, z["ae_J"])
    assert q.tobytes() == z["ae_q"].tobytes()


@pytest.mark.parametrize("meta", CASES, ids=[c["name"] for c in CASES])
def test_replay_walks_bit_exact(meta):
    z, g = load_case(meta)
    an = (z["an_J"], z["an_q"])
    ae = (z["ae_off"], z["ae_J"], z["ae_q"])
    walks, pos = O.simulate_walks_replay(g, an, ae, meta["walk_length"], z["starts"], z["uniforms"])
    assert pos == meta["n_uniforms"]
    for i, w in enumerate(walks):
        assert len(w) == z["lens"][i]
        assert w == z["walks"][i, :len(w)].tolist()


def test_survey_walk_kat():
    meta = [c for c in CASES if c["name"] == "karate_p025_q4"][0]
    assert meta["first_walk_ids"] == [12, 1, 12, 1, 12, 1, 12, 1, 3, 1, 3, 14, 3, 2, 3, 2]
    assert meta["sha256_alias_edges"] == "684d12d03f94be900252aedc075fc80ae1b5fa3963f4d1c90e28c2338bcba64e"
