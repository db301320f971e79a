"""Partition / shared-node index maps — vectorised, sequence-exact with the reference.

Every function returns exactly the sequence the corresponding list-scanning function of
/root/reference/Tools/Distributed_tools.py returns (as an int64 array instead of a Python list),
in O(n log n) instead of O(n^2):

    rankwise_dist        Distributed_tools.py:14-24   elements of a rank + nodes in first-appearance order
    find_shared_nodes    Distributed_tools.py:29-40   shared nodes, other-rank-major order, de-duplicated
    sort_shared          Distributed_tools.py:44-51   sorted union of all ranks' shared lists
    Dirichlet_rank_dist  Distributed_tools.py:55-62   local clamped DOFs, ascending local position
    local_mat_node       Distributed_tools.py:66-73   positions of global ids in a local list
    node_to_dof          commons.py:66-71             interleaved DOF numbering d*g+i

On top of these, `halo_plan` derives what the exchange kernels need: for each pair of ranks the
common nodes in a canonical (ascending global id) order, and for every shared DOF the holders in
ascending rank order — the association in which syn_cpus (Distributed_tools.py:83-86) adds the
partial forces.
"""
from __future__ import annotations

import numpy as np


def _first_appearance_unique(seq):
    seq = np.asarray(seq, dtype=np.int64).ravel()
    if seq.size == 0:
        return seq
    uniq, first = np.unique(seq, return_index=True)
    return uniq[np.argsort(first, kind="stable")]


def node_to_dof(d, ls, P):
    """commons.py:66-71 — [d*g + i for g in P for i in ls] as int64."""
    P = np.asarray(P, dtype=np.int64).ravel()
    ls = np.asarray(ls, dtype=np.int64).ravel()
    return (d * P[:, None] + ls[None, :]).ravel()


def rankwise_dist(rank, epart, cells):
    """Distributed_tools.py:14-24 — (ascending element ids of `rank`, their nodes in first-appearance order)."""
    ele = np.nonzero(np.asarray(epart) == rank)[0].astype(np.int64)
    nodes = _first_appearance_unique(np.asarray(cells)[ele])
    return ele, nodes


def find_shared_nodes(rank, size, rank_nodal_list):
    """Distributed_tools.py:29-40 — for r != rank ascending, r's nodes (in r's order) that `rank` also
    holds, each listed once at its first occurrence."""
    mine = np.asarray(rank_nodal_list[rank], dtype=np.int64)
    others = [np.asarray(rank_nodal_list[r], dtype=np.int64) for r in range(size) if r != rank]
    if not others:
        return np.zeros(0, dtype=np.int64)
    cat = np.concatenate(others) if others else np.zeros(0, dtype=np.int64)
    cat = cat[np.isin(cat, mine)]
    return _first_appearance_unique(cat)


def sort_shared(shared_lists):
    """Distributed_tools.py:44-51 — np.sort of the de-duplicated union."""
    nonempty = [np.asarray(s, dtype=np.int64) for s in shared_lists if len(s)]
    if not nonempty:
        return np.zeros(0, dtype=np.int64)
    return np.unique(np.concatenate(nonempty))


def local_mat_node(G_ID, L_N):
    """Distributed_tools.py:66-73 — position in L_N of each g in G_ID (ids absent from L_N are skipped)."""
    G = np.asarray(G_ID, dtype=np.int64).ravel()
    L = np.asarray(L_N, dtype=np.int64).ravel()
    if G.size == 0 or L.size == 0:
        return np.zeros(0, dtype=np.int64)
    order = np.argsort(L, kind="stable")
    pos = np.searchsorted(L[order], G)
    pos = np.clip(pos, 0, L.size - 1)
    hit = L[order][pos] == G
    return order[pos][hit].astype(np.int64)


def Dirichlet_rank_dist(D_node, Local_N_list):
    """Distributed_tools.py:55-62 — DOFs [3k,3k+1,3k+2] of ascending local positions k whose node is clamped."""
    L = np.asarray(Local_N_list, dtype=np.int64)
    k = np.nonzero(np.isin(L, np.asarray(D_node, dtype=np.int64)))[0]
    return node_to_dof(3, [0, 1, 2], k)


def partition_maps(epart, cells, size, dirichlet=None):
    """All per-rank maps of Data_prepare.py:104-144 for a `size`-way element partition."""
    per = []
    for r in range(size):
        ele, nodes = rankwise_dist(r, epart, cells)
        per.append(dict(rank=r, Local_ele_list=ele, Local_nodal_list=nodes))
    lists = [p["Local_nodal_list"] for p in per]
    for p in per:
        p["shared_nodes"] = find_shared_nodes(p["rank"], size, lists)
        p["loc_dof_shared"] = node_to_dof(3, [0, 1, 2], local_mat_node(p["shared_nodes"], p["Local_nodal_list"]))
        if dirichlet is not None:
            p["Local_Dirichlet"] = Dirichlet_rank_dist(dirichlet, p["Local_nodal_list"])
    return per, sort_shared([p["shared_nodes"] for p in per])


def halo_plan(rank, size, rank_nodal_list):
    """Exchange description for `rank`, derived from the same lists syn_cpus sees.

    Returns a dict with
      shared_pos   (s,)  local positions (into Local_nodal_list) of this rank's shared nodes, ascending
                         global node id (the canonical interface order both sides of an interface agree on)
      neighbours   list of ranks sharing at least one node, ascending
      send_idx[nb] positions into `shared_pos` of the nodes shared with `nb`, ascending global id —
                   rank `nb` builds the same list of global ids, so message k of one side is message k
                   of the other
      holders_ptr, holders_rank  CSR over the s shared nodes: ranks holding the node, ASCENDING — the
                   order in which Distributed_tools.py:85-86 accumulates (`for i in range(size)`)
      holders_slot for each (node, holder): -1 if holder == rank (own partial force) else the index of
                   that node inside the message received from that holder
    """
    mine = np.asarray(rank_nodal_list[rank], dtype=np.int64)
    order = np.argsort(mine, kind="stable")
    mine_sorted = mine[order]
    nb_nodes = {}
    for r in range(size):
        if r == rank:
            continue
        other = np.asarray(rank_nodal_list[r], dtype=np.int64)
        common = np.intersect1d(mine, other, assume_unique=True)   # ascending global id
        if common.size:
            nb_nodes[r] = common
    neighbours = sorted(nb_nodes)
    if neighbours:
        shared_sorted = np.unique(np.concatenate([nb_nodes[r] for r in neighbours]))
    else:
        shared_sorted = np.zeros(0, dtype=np.int64)
    shared_pos = order[np.searchsorted(mine_sorted, shared_sorted)].astype(np.int64)
    send_idx = {r: np.searchsorted(shared_sorted, nb_nodes[r]).astype(np.int64) for r in neighbours}
    s = shared_sorted.size
    # holders of each shared node in ascending rank order (own rank included)
    cnt = np.ones(s, dtype=np.int64)
    for r in neighbours:
        cnt[send_idx[r]] += 1
    ptr = np.zeros(s + 1, dtype=np.int64)
    np.cumsum(cnt, out=ptr[1:])
    h_rank = np.empty(ptr[-1], dtype=np.int64)
    h_slot = np.empty(ptr[-1], dtype=np.int64)
    fill = ptr[:-1].copy()
    for r in sorted(neighbours + [rank]):
        if r == rank:
            idx = np.arange(s)
            h_rank[fill[idx]] = r
            h_slot[fill[idx]] = -1
            fill[idx] += 1
        else:
            idx = send_idx[r]
            h_rank[fill[idx]] = r
            h_slot[fill[idx]] = np.arange(idx.size)
            fill[idx] += 1
    return dict(shared_nodes_sorted=shared_sorted, shared_pos=shared_pos, neighbours=neighbours,
                send_idx=send_idx, holders_ptr=ptr, holders_rank=h_rank, holders_slot=h_slot)
