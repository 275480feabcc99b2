"""CPU property tests of two pieces of pure host logic:
  * the chunk-major gathered-table layout of the row partition (RowPartition.table_row / chunk_pieces / _remap_ /
    bounds): a bijection onto [0, G*n_pad), chunk c of all ranks contiguous, consistent with the in-place remap;
  * the DGL stand-in under tests/fake_dgl that lets the unmodified reference layer run: the documented DGL semantics
    it models (zero-fill of nodes without in-edges for every builtin reducer, local_scope restoring the frames,
    edge-id order of the gathered views, multigraph / self-loop counting) — the pin's own pins.
"""
import os
import sys

import pytest
import torch

import sirgcn_b200  # noqa: F401
from sirgcn_b200 import partition
from sirgcn_b200.graph import CompressedRows

FAKE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fake_dgl")
if FAKE not in sys.path:
    sys.path.insert(0, FAKE)
import dgl  # noqa: E402
from dgl import function as fn  # noqa: E402


def _part(n, world, rank, layout):
    e = torch.empty(0, dtype=torch.int32)
    n_pad, lo, hi = partition.RowPartition.bounds(n, rank, world, layout)
    ip = torch.zeros(hi - lo + 1, dtype=torch.int32)
    ones = torch.ones(world * n_pad)
    return partition.RowPartition(n, rank, world, CompressedRows(ip, e, None), CompressedRows(ip.clone(), e.clone(), None),
                                  ones, ones.clone(), ones.clone(), layout_chunks=layout)


@pytest.mark.parametrize("n,world,layout", [(37, 2, 1), (37, 2, 4), (37, 3, 8), (1000, 8, 4), (1001, 8, 4), (5, 8, 4),
                                            (64, 4, 16), (100_003, 8, 4)])
def test_chunk_major_layout_is_a_bijection_with_contiguous_chunks(n, world, layout):
    parts = [_part(n, world, r, layout) for r in range(world)]
    p0 = parts[0]
    G, n_pad, step = world, p0.n_pad, p0.step
    assert n_pad % layout == 0 and G * n_pad >= n and step * layout == n_pad
    # every rank sees the same padded size; the ranks' row ranges tile [0, n)
    assert all(p.n_pad == n_pad for p in parts)
    assert [p.lo for p in parts] == [min(n, r * n_pad) for r in range(G)] and parts[-1].hi == n
    rows = torch.tensor([[p0.table_row(r, i) for i in range(n_pad)] for r in range(G)])
    assert sorted(rows.flatten().tolist()) == list(range(G * n_pad))                     # a bijection
    for c in range(layout):                                                              # chunk c of all ranks: one block
        blk = rows[:, c * step:(c + 1) * step].flatten()
        assert blk.min() == c * G * step and blk.max() == (c + 1) * G * step - 1
        for r in range(G):                                                               # rank-major inside the block
            assert rows[r, c * step:(c + 1) * step].tolist() == list(range((c * G + r) * step, (c * G + r + 1) * step))
    # the in-place remap of global ids agrees with table_row
    g = torch.arange(G * n_pad, dtype=torch.int32)
    if layout > 1 and world > 1:
        p0._remap_(g, block=97)
    assert g.tolist() == [p0.table_row(i // n_pad, i % n_pad) for i in range(G * n_pad)]
    # chunk_pieces cuts any row range at the layout-chunk boundaries and covers it exactly once
    for lo, hi in [(0, n_pad), (0, 0), (1, max(1, n_pad - 1)), (step // 2, min(n_pad, step + step // 2 + 1))]:
        pieces = p0.chunk_pieces(lo, hi)
        assert [a for a, _, _ in pieces] == ([lo] + [b for _, b, _ in pieces[:-1]] if pieces else [])
        assert (pieces[-1][1] if pieces else lo) == max(lo, hi) or hi <= lo
        for a, b, c in pieces:
            assert c * step <= a < b <= (c + 1) * step


def test_walk_mode_switches_to_plain_grids_only_above_one_rank(monkeypatch):
    from sirgcn_b200 import function
    p1, p2 = _part(10, 1, 0, 1), _part(10, 2, 0, 1)
    assert function.WALK_PERSISTENT is True
    with partition._walk_mode(p1):
        assert function.WALK_PERSISTENT is True
    with partition._walk_mode(p2):
        assert function.WALK_PERSISTENT is False
        with partition._walk_mode(p1):                       # nested single-rank scope keeps what it found
            assert function.WALK_PERSISTENT is False
    assert function.WALK_PERSISTENT is True
    monkeypatch.setenv("SIRGCN_PARTITION_PERSIST", "1")
    with partition._walk_mode(p2):
        assert function.WALK_PERSISTENT is True
    with pytest.raises(RuntimeError):                        # restored even when the body raises
        with partition._walk_mode(_part(10, 2, 0, 1)):
            raise RuntimeError("x")
    assert function.WALK_PERSISTENT is True


# ---- the DGL stand-in ---------------------------------------------------------------------------------------
def test_standin_reducers_zero_fill_and_count_multi_edges():
    # edges by id: 0:2->1  1:0->1  2:2->1 (duplicate)  3:1->1 (self loop)  4:0->3 ; node 4 has no in-edges, node 2 none either
    g = dgl.graph((torch.tensor([2, 0, 2, 1, 0]), torch.tensor([1, 1, 1, 1, 3])), num_nodes=5)
    assert g.in_degrees().tolist() == [0, 4, 0, 1, 0] and g.out_degrees().tolist() == [2, 1, 2, 0, 0]
    g.ndata["h"] = torch.tensor([[1.0], [10.0], [100.0], [1000.0], [-5.0]])
    g.edata["w"] = torch.tensor([[1.0], [2.0], [3.0], [4.0], [5.0]])
    seen = {}

    def msg(edges):
        seen["src"], seen["dst"], seen["w"] = edges.src["h"].clone(), edges.dst["h"].clone(), edges.data["w"].clone()
        return {"m": edges.src["h"] * edges.data["w"] - 1000.0}

    for name, want in (("sum", [0.0, 100 + 2 + 300 + 40 - 4000, 0.0, 5 - 1000.0, 0.0]),
                       ("mean", [0.0, (100 + 2 + 300 + 40 - 4000) / 4, 0.0, 5 - 1000.0, 0.0]),
                       ("max", [0.0, 300 - 1000.0, 0.0, 5 - 1000.0, 0.0]),           # NOT max(., 0): zero only when empty
                       ("min", [0.0, 2 - 1000.0, 0.0, 5 - 1000.0, 0.0])):
        with g.local_scope():
            g.update_all(msg, getattr(fn, name)("m", "ft"))
            assert g.ndata["ft"].squeeze(1).tolist() == pytest.approx(want), name
        assert set(g.ndata) == {"h"} and set(g.edata) == {"w"}                       # local_scope restored the frames
    # the message UDF saw the endpoint features gathered in EDGE-ID order
    assert seen["src"].squeeze(1).tolist() == [100.0, 1.0, 100.0, 10.0, 1.0]
    assert seen["dst"].squeeze(1).tolist() == [10.0, 10.0, 10.0, 10.0, 1000.0]
    assert seen["w"].squeeze(1).tolist() == [1.0, 2.0, 3.0, 4.0, 5.0]


def test_standin_rejects_what_it_does_not_model():
    g = dgl.graph((torch.tensor([0]), torch.tensor([1])), num_nodes=2)
    with pytest.raises(ValueError):
        g.ndata["x"] = torch.zeros(3, 1)                       # wrong number of rows (DGL raises too)
    with pytest.raises(NotImplementedError):
        g.update_all("not a udf", fn.sum("m", "ft"))
    with pytest.raises(ValueError):
        dgl.graph((torch.tensor([0]), torch.tensor([5])), num_nodes=2)
    with pytest.raises(NotImplementedError):
        g.edges(form="all")
    b = dgl.batch([g, g])
    assert b.batch_size == 2 and b.num_nodes() == 4 and b.edges()[0].tolist() == [0, 2] and b.edges()[1].tolist() == [1, 3]
