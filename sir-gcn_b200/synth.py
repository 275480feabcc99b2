"""Synthetic graphs of the shapes named in BASELINE.json `configs` (there is no network for the real
datasets).  Pure torch index arithmetic, usable on CPU (tests, CPU baseline) and on the GPU (bench;
the 2 B-edge graph cannot even be held by the host).  Every generator is deterministic in `seed`.

Shapes follow SURVEY.md §8(d):
  Z  zinc_like     128 molecules x 23 atoms, 25 bonds in both directions (50 directed edges), bond types
  A  arxiv_like    169,343 nodes / 1,166,243 edges, power-law in-degree (exponent 2.3, capped)
  C  cifar_like    128 super-pixel graphs, n ~ U{85..150}, directed kNN k = 8
  P  powerlaw      N nodes / E edges, power-law in-degree, uniform sources, emitted destination-sorted
"""
from __future__ import annotations

import torch


def _gen(seed, device):
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    return g


def zinc_like(num_graphs=128, nodes=23, extra_bonds=3, num_atom_types=28, num_bond_types=4, seed=0, device="cpu"):
    """returns src, dst (int64 [E]), num_nodes, atom_type [N], bond_type [E]"""
    g = _gen(seed, "cpu")
    srcs, dsts, bts = [], [], []
    for b in range(num_graphs):
        base = b * nodes
        bonds = set()
        for i in range(1, nodes):                       # random spanning tree
            j = int(torch.randint(0, i, (1,), generator=g))
            bonds.add((j, i))
        while len(bonds) < nodes - 1 + extra_bonds:     # ring-closing bonds, no self loops / duplicates
            i, j = (int(x) for x in torch.randint(0, nodes, (2,), generator=g))
            if i != j and (min(i, j), max(i, j)) not in bonds:
                bonds.add((min(i, j), max(i, j)))
        bonds = sorted(bonds)
        bt = torch.randint(0, num_bond_types, (len(bonds),), generator=g)
        u = torch.tensor([p[0] for p in bonds]) + base
        v = torch.tensor([p[1] for p in bonds]) + base
        srcs += [u, v]
        dsts += [v, u]
        bts += [bt, bt]
    src, dst, bond = torch.cat(srcs), torch.cat(dsts), torch.cat(bts)
    n = num_graphs * nodes
    atom = torch.randint(0, num_atom_types, (n,), generator=g)
    return src.to(device), dst.to(device), n, atom.to(device), bond.to(device)


def _powerlaw_degrees(n, e, alpha, max_deg, gen, device):
    """integer degrees with Pareto(alpha) tail, capped, summing to exactly e"""
    u = torch.rand(n, generator=gen, device=device, dtype=torch.float64).clamp_(min=1e-12)
    w = u.pow_(-1.0 / (alpha - 1.0)).clamp_(max=float(max_deg))
    w.mul_(e / float(w.sum()))
    deg = w.floor().to(torch.int64)
    rem = int(e - int(deg.sum()))
    if rem > 0:
        deg[:rem] += 1
    elif rem < 0:                                        # only if the cap made rounding overshoot
        idx = torch.nonzero(deg > 0).flatten()[:-rem]
        deg[idx] -= 1
    return deg


def powerlaw(num_nodes, num_edges, alpha=2.3, max_deg=None, seed=0, device="cpu", index_dtype=torch.int32):
    """Chung-Lu style: in-degree ~ power law (exponent alpha), sources uniform; multi-edges and
    self-loops kept.  Returns (src, dst, num_nodes) already destination-sorted."""
    gen = _gen(seed, device)
    if max_deg is None:
        max_deg = max(64, num_nodes // 4)
    deg = _powerlaw_degrees(num_nodes, num_edges, alpha, max_deg, gen, device)
    dst = torch.repeat_interleave(torch.arange(num_nodes, dtype=index_dtype, device=device), deg,
                                  output_size=int(num_edges))
    del deg
    src = torch.empty(num_edges, dtype=index_dtype, device=device)
    step = 1 << 28                                       # bounded temporaries for the 2 B-edge case
    for lo in range(0, num_edges, step):
        hi = min(num_edges, lo + step)
        src[lo:hi] = torch.randint(0, num_nodes, (hi - lo,), generator=gen, device=device, dtype=index_dtype)
    return src, dst, num_nodes


def arxiv_like(num_nodes=169_343, num_edges=1_166_243, seed=0, device="cpu", bidirected_self_loops=False):
    """power-law in-degree capped at ~13 k (ogbn-arxiv's max in-degree); edge ids shuffled so the COO
    is NOT pre-sorted (the builder must do real work)."""
    src, dst, n = powerlaw(num_nodes, num_edges, alpha=2.3, max_deg=13_000, seed=seed, device=device,
                           index_dtype=torch.int64)
    perm = torch.randperm(num_edges, generator=_gen(seed + 1, device), device=device)
    src, dst = src[perm], dst[perm]
    if bidirected_self_loops:                            # published recipe (ogbn-arxiv/README.md:21)
        loops = torch.arange(n, device=device)
        src, dst = torch.cat([src, dst, loops]), torch.cat([dst, src, loops])
    return src, dst, n


def cifar_like(num_graphs=128, lo=85, hi=150, k=8, seed=0, device="cpu"):
    """directed kNN graphs on uniform 2-D points: every node has in-degree exactly k, no self loops.
    Edge direction neighbour -> node (DGL's knn_graph convention). Returns src, dst, num_nodes, pos, dist."""
    g = _gen(seed, "cpu")
    sizes = torch.randint(lo, hi + 1, (num_graphs,), generator=g)
    srcs, dsts, poss, dists = [], [], [], []
    base = 0
    for n in sizes.tolist():
        pos = torch.rand(n, 2, generator=g)
        dmat = torch.cdist(pos, pos)
        dmat.fill_diagonal_(float("inf"))
        dist, nbr = dmat.topk(k, dim=1, largest=False)
        dsts.append(torch.arange(n).repeat_interleave(k) + base)
        srcs.append(nbr.flatten() + base)
        dists.append(dist.flatten())
        poss.append(pos)
        base += n
    return (torch.cat(srcs).to(device), torch.cat(dsts).to(device), base,
            torch.cat(poss).to(device), torch.cat(dists).to(device))
