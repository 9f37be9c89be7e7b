"""Pins the oracle restatement (oracle/oracle_np.py + oracle/oracle_c.c) against the golden
fixtures written by the UNMODIFIED reference: derived adjacencies, edge star, stage
intermediates and the complete output mesh of every pass. Integers bit-exact, reals 1e-12."""
import os
import sys

import numpy as np
import pytest

import parity
from conftest import golden_files

sys.path.insert(0, os.path.join(parity.ROOT, "oracle"))
import oracle_np as onp  # noqa: E402


@pytest.mark.parametrize("path", golden_files(), ids=lambda p: os.path.basename(p).split(".")[0])
def test_oracle_matches_reference(path):
    fx = parity.load(path)
    m = onp.mesh_from_fixture(fx)
    rep = parity.Report()
    dim = m.dim
    for hd in range(2, dim + 1):
        rep.eq("verts_of%d" % hd, m.verts_of(hd), fx["in:verts_of%d" % hd])
    if dim == 3:
        ab2b, codes = m.ask_down(3, 1)
        rep.eq("down31", ab2b, fx["in:down31"])
        rep.eq("codes31", codes, fx["in:codes31"])
    for lo in range(dim):
        for hi in range(lo + 1, dim + 1):
            a2ab, ab2b, codes = m.ask_up(lo, hi)
            s = "in:up%d%d" % (lo, hi)
            rep.eq(s + ":a2ab", a2ab, fx[s + ":a2ab"])
            rep.eq(s + ":ab2b", ab2b, fx[s + ":ab2b"])
            rep.eq(s + ":codes", codes, fx[s + ":codes"])
    xadj, adj = m.ask_star_edges()
    rep.eq("star:a2ab", xadj, fx["in:star1:a2ab"])
    rep.eq("star:ab2b", adj, fx["in:star1:ab2b"])
    new, info = onp.refine_by_size(m, float(fx["opts:max_length_desired"][0]), float(fx["opts:min_quality_allowed"][0]))
    rep.eq("candidate", info["candidate"], fx["mid:candidate"])
    for k in ("cands2edges", "key", "rep_vertex2md_order"):
        if "mid:" + k in fx:
            rep.eq(k, info[k], fx["mid:" + k])
    for k in ("mident_metrics", "cand_quals"):
        if "mid:" + k in fx:
            rep.close(k, info[k], fx["mid:" + k])
    rep.eq("did", int(new is not None), int(fx["did"][0]))
    if new is not None:
        for d in range(dim + 1):
            rep.eq("nents%d" % d, new.nents[d], int(fx["out:nents%d" % d][0]))
        for d in range(1, dim + 1):
            rep.eq("down%d" % d, new.down[d][0], fx["out:down%d" % d])
            if d > 1:
                rep.eq("codes%d" % d, new.down[d][1], fx["out:codes%d" % d])
        for d in range(dim + 1):
            tp = "out:tag%d:" % d
            for k in fx:
                if k.startswith(tp) and not k.endswith(":ncomps"):
                    name = k[len(tp):]
                    if name not in new.tags[d]:
                        rep.fail.append("oracle lacks tag " + k)
                    elif fx[k].dtype == np.float64:
                        rep.close(k, new.get(d, name), fx[k])
                    else:
                        rep.eq(k, new.get(d, name), fx[k])
    rep.assert_ok()


def test_oracle_find_unique_and_reflect_down_known_answers():
    # src/unit_mesh.cpp:129-209
    assert onp.find_unique(np.array([0, 1, 2, 2, 3, 0], dtype=np.int32), 2, 1).tolist() == [0, 1, 0, 2, 3, 0, 1, 2, 2, 3]
    l, c = onp.reflect_down(np.array([0, 1, 2, 3], dtype=np.int32),
                            np.array([0, 1, 2, 0, 3, 1, 1, 3, 2, 2, 3, 0], dtype=np.int32), 3, 2)
    assert l.tolist() == [0, 1, 2, 3] and c.tolist() == [1, 1, 1, 1]
    l, c = onp.reflect_down(np.array([0, 1, 2, 2, 3, 0], dtype=np.int32),
                            np.array([0, 1, 1, 2, 2, 3, 3, 0, 0, 2], dtype=np.int32), 2, 1)
    assert l.tolist() == [0, 1, 4, 2, 3, 4]
