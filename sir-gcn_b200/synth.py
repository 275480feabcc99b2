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


# ---- counter-based variant: any rank can enumerate any slice of the SAME global graph --------------------
def _hash_src(e, num_nodes, seed):
    """source node of global edge number e (int64 tensor, e < 2^32): murmur-style 32-bit finalizer, then a
    multiply-shift onto [0, num_nodes); a pure function of (e, seed), so every rank of a partitioned run sees
    the same graph without exchanging edges"""
    m = 0xFFFFFFFF
    x = (e ^ (int(seed) * 0x9E3779B1 & m)) & m
    x = ((x ^ (x >> 16)) * 0x45D9F3B) & m
    x = ((x ^ (x >> 16)) * 0x45D9F3B) & m
    x = x ^ (x >> 16)
    return (x * int(num_nodes)) >> 32


def powerlaw_indptr(num_nodes, num_edges, alpha=2.3, max_deg=None, seed=0, device="cpu"):
    """global in-CSR row pointer (int64 [N+1]) of the power-law graph; identical on every rank"""
    gen = _gen(seed, device)
    if max_deg is None:
        max_deg = max(64, num_nodes // 4)
    deg = _powerlaw_degrees(num_nodes, num_edges, alpha, max_deg, gen, device)
    indptr = torch.zeros(num_nodes + 1, dtype=torch.int64, device=device)
    torch.cumsum(deg, 0, out=indptr[1:])
    return indptr


def powerlaw_hashed_rows(indptr, num_nodes, lo, hi, seed=0, index_dtype=torch.int32, step=1 << 27, want_dst=True):
    """edges whose DESTINATION lies in [lo, hi) of the hashed power-law graph, destination-sorted:
    (src global ids, dst global ids) — the in-CSR slice of a destination-row partition"""
    dev = indptr.device
    e_lo, e_hi = int(indptr[lo]), int(indptr[hi])
    n = e_hi - e_lo
    src = torch.empty(n, dtype=index_dtype, device=dev)
    for a in range(e_lo, e_hi, step):
        b = min(e_hi, a + step)
        src[a - e_lo:b - e_lo] = _hash_src(torch.arange(a, b, dtype=torch.int64, device=dev), num_nodes, seed).to(index_dtype)
    if not want_dst:
        return src, None
    deg = indptr[lo + 1:hi + 1] - indptr[lo:hi]
    dst = torch.repeat_interleave(torch.arange(lo, hi, dtype=index_dtype, device=dev), deg, output_size=n)
    return src, dst


def powerlaw_hashed_cols(indptr, num_nodes, lo, hi, seed=0, index_dtype=torch.int32, step=1 << 27):
    """edges whose SOURCE lies in [lo, hi), in global edge order (i.e. destination-sorted): (src, dst) global
    ids — the out-CSC slice of a destination-row partition; found by scanning the hash of every edge number"""
    dev = indptr.device
    num_edges = int(indptr[-1])
    srcs, dsts = [], []
    for a in range(0, num_edges, step):
        b = min(num_edges, a + step)
        e = torch.arange(a, b, dtype=torch.int64, device=dev)
        s = _hash_src(e, num_nodes, seed)
        keep = (s >= lo) & (s < hi)
        e, s = e[keep], s[keep]
        d = torch.searchsorted(indptr, e, right=True) - 1
        srcs.append(s.to(index_dtype))
        dsts.append(d.to(index_dtype))
        del e, s, d, keep
    return torch.cat(srcs), torch.cat(dsts)


def powerlaw_hashed(num_nodes, num_edges, alpha=2.3, max_deg=None, seed=0, device="cpu", index_dtype=torch.int32):
    """the whole hashed power-law graph as destination-sorted COO (src, dst, num_nodes)"""
    indptr = powerlaw_indptr(num_nodes, num_edges, alpha, max_deg, seed, device)
    src, dst = powerlaw_hashed_rows(indptr, num_nodes, 0, num_nodes, seed, index_dtype)
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
