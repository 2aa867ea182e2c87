"""Literal (quadratic) restatement of the reference's contact-surface bookkeeping — TEST INFRASTRUCTURE ONLY.

get_surface_triangle (HAKAI-v0.0.2/Julia/HAKAI_j.jl:1996-2164), add_surface_triangle (J2:2167-2245) and the update
loop after element deletion (J2:767-804), as the same nested scans the Julia code performs (only the innermost
row comparison is a NumPy expression).  Used by tests/test_oracle_anchors.py to check the sort-based production
versions (hakai_fem_b200/model_setup.py, hk_engine.cu `update_surfaces`) and the C++ oracle against a second reading
of the reference, including its quirks: the scan `j = 1 : 6nE-1` never emits the last face, duplicates are paired
first-with-next, a self-contact pair only receives nodes (the `if / elseif` of J2:784-797).
"""
import numpy as np


def get_surface_triangle(surfaces, surfaces_eleid, contact_element=None):
    """surfaces (F,4) 1-based part-local faces of ALL elements of the instance (array_element = 1:nE)."""
    F = surfaces.shape[0]
    srt = np.sort(surfaces, axis=1)
    dup = set()
    c_surf, c_ele = [], []
    for j in range(F - 1):                                      # J2:2040: the last row is never visited
        if j in dup:
            continue
        same = np.flatnonzero(np.all(srt[j + 1:] == srt[j], axis=1))
        if len(same):                                           # first later row with the same node set
            dup.add(j + 1 + int(same[0]))
            continue
        c_surf.append(surfaces[j])
        c_ele.append(surfaces_eleid[j])
    c_surf = np.array(c_surf, np.int64).reshape(-1, 4)
    c_ele = np.array(c_ele, np.int64)
    if contact_element is not None:                             # J2:2094-2119
        keep = np.array([e in set(int(v) for v in contact_element) for e in c_ele], bool)
        c_surf, c_ele = c_surf[keep], c_ele[keep]
    tri = np.zeros((2 * len(c_surf), 3), np.int64)
    tri[0::2] = c_surf[:, [0, 1, 2]]
    tri[1::2] = c_surf[:, [2, 3, 0]]
    return tri, np.repeat(c_ele, 2), np.unique(tri)


def add_surface_triangle(surfaces, surfaces_eleid, ele_id):
    """Faces of OTHER elements that share a node set with one of the 6 faces of element `ele_id` (1-based)."""
    srt = np.sort(surfaces, axis=1)
    other = surfaces_eleid != ele_id
    add_s, add_e = [], []
    for j in range(6):
        sj = srt[6 * (ele_id - 1) + j]
        hit = np.flatnonzero(other & np.all(srt == sj, axis=1))
        if len(hit):                                            # first in face order, then `break`
            add_s.append(surfaces[hit[0]])
            add_e.append(surfaces_eleid[hit[0]])
    add_s = np.array(add_s, np.int64).reshape(-1, 4)
    tri = np.zeros((2 * len(add_s), 3), np.int64)
    tri[0::2] = add_s[:, [0, 1, 2]]
    tri[1::2] = add_s[:, [2, 3, 0]]
    return tri, np.repeat(np.array(add_e, np.int64), 2), np.unique(tri)


def _append_unique(lst, new):
    seen = set(lst)
    for v in new:
        if int(v) not in seen:
            seen.add(int(v))
            lst.append(int(v))


def replay_deletions(instances, element_instance, pairs, deleted):
    """J2:767-804 for the global element ids `deleted` (in deletion order).
    instances: list of dicts {surfaces, surfaces_eleid, node_offset, element_offset};
    pairs: list of dicts {i_instance, j_instance, c_nodes_i, c_nodes_j, c_triangles (n,3), c_triangles_eleid}
    (lists / arrays, 1-based global ids) — updated in place."""
    for p in pairs:
        p["c_nodes_i"], p["c_nodes_j"] = list(map(int, p["c_nodes_i"])), list(map(int, p["c_nodes_j"]))
        p["c_triangles"] = [tuple(map(int, r)) for r in np.asarray(p["c_triangles"]).reshape(-1, 3)]
        p["c_triangles_eleid"] = list(map(int, p["c_triangles_eleid"]))
    for g in deleted:
        inst_id = int(element_instance[g - 1])
        I = instances[inst_id - 1]
        tri, tele, nodes = add_surface_triangle(I["surfaces"], I["surfaces_eleid"], int(g) - I["element_offset"])
        for p in pairs:
            if p["i_instance"] == inst_id:
                _append_unique(p["c_nodes_i"], nodes + I["node_offset"])
            elif p["j_instance"] == inst_id:
                _append_unique(p["c_nodes_j"], nodes + I["node_offset"])
                p["c_triangles_eleid"].extend(int(v) + I["element_offset"] for v in tele)
                p["c_triangles"].extend(tuple(int(v) + I["node_offset"] for v in r) for r in tri)
    return pairs
