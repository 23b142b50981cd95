#!/usr/bin/env python
"""Generates feonet_navier_stokes_b200/csrc/feo_lattice_gen.inc: the macro-stencil of the lattice plan (feo_lattice.h).

The lattice plan evaluates the fused residual on a structured "right-diagonal" P2-P1 mesh (the mesh family of the reference's
`RectangleMesh` set-ups, `FEONet_steady_Navier-Stokes/assemble_fenics.py:50-56`, and of BASELINE.json configs[4]) cell by
cell: a cell = the lattice nodes (2ci+{0,1}, 2cj+{0,1}) -- a vertex node V, its right / upper / diagonal edge nodes H, T, D --
plus the pressure dof of V; all its couplings lie in the 5 x 5 node window [-2, 2]^2 around V.  WHICH couplings can be non-zero
(per matrix A, B1, B2) is a property of the element family and the forms, not of the mesh size or the boundary conditions;
this script derives that union pattern from small operators assembled by the repo's own fixture assembler (all cell classes:
interior, edges, corners, Dirichlet and natural boundaries) and emits

  * the coefficient-table layout (one descriptor per table entry: which matrix entry it holds), used by the host planner
    (feo_lattice_plan.cpp) to fill one table per cell class from the CSR handed to feo_op_create and to VERIFY that the CSR is
    fully explained by the pattern (otherwise the tile plan is used);
  * the straight-line bodies of the forward (row-owned) and backward (column-owned) cell evaluation as macro lists, compiled
    into the CUDA kernels (feo_lattice.cu: coefficients are constant-bank operands) and into the fp64 host replay.

Run:  python tools/gen_lattice_stencil.py   (re-generates the .inc; the output is committed)
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from feonet_navier_stokes_b200.fixtures.taylor_hood import assemble_operators, structured_mesh  # noqa: E402

OUT = os.path.join(ROOT, "feonet_navier_stokes_b200", "csrc", "feo_lattice_gen.inc")
TARGETS = [(0, 0), (1, 0), (0, 1), (1, 1)]  # V, H, T, D relative to the cell origin (2ci, 2cj)
MATS = ("A", "B1", "B2")


def cell_entries(fx, n):
    """{(mat, row(dx,dy,comp), col(dx,dy,comp)): value} of every cell's rows, in window coordinates."""
    m = 2 * n + 1
    mats = {}
    for name in MATS:
        M = getattr(fx, name).tocsr().astype(np.float32)
        M.eliminate_zeros()
        M.sort_indices()
        mats[name] = M
    iu1, iu2, ip = fx.idx_u1, fx.idx_u2, fx.idx_p
    p2_to_p1 = -np.ones(m * m, dtype=np.int64)
    p2_to_p1[fx.mesh.p1_to_p2] = np.arange((n + 1) ** 2)

    def dof(x, y, comp):
        if x < 0 or y < 0 or x >= m or y >= m:
            return -1
        k = y * m + x
        if comp == 0:
            return int(iu1[k])
        if comp == 1:
            return int(iu2[k])
        return int(ip[p2_to_p1[k]]) if (x % 2 == 0 and y % 2 == 0) else -1

    out = {}
    for cj in range(n + 1):
        for ci in range(n + 1):
            ox, oy = 2 * ci, 2 * cj
            where = {}
            for dx in range(-2, 3):
                for dy in range(-2, 3):
                    for cc in (0, 1, 2):
                        d = dof(ox + dx, oy + dy, cc)
                        if d >= 0:
                            where[d] = (dx, dy, cc)
            ent = {}
            for (tx, ty) in TARGETS:
                for tc in (0, 1, 2):
                    r = dof(ox + tx, oy + ty, tc)
                    if r < 0:
                        continue
                    for name in MATS:
                        if name != "A" and tc == 2:
                            continue  # B1 / B2 rows of pressure dofs never enter the residual (convection is zero there)
                        M = mats[name]
                        for c, v in zip(M.indices[M.indptr[r]:M.indptr[r + 1]], M.data[M.indptr[r]:M.indptr[r + 1]]):
                            assert c in where, "entry outside the 5x5 window"
                            ent[(name, (tx, ty, tc), where[c])] = float(v)
            out[(ci, cj)] = ent
    return out


def union_pattern():
    keys = set()
    for n, bc in ((6, "channel_flow"), (7, "channel_flow"), (6, "lower"), (6, "none")):
        fx = assemble_operators(structured_mesh(n), mu=0.1, bc=bc, ordering="interleaved", with_convection=True)
        for ent in cell_entries(fx, n).values():
            keys |= set(ent.keys())
    # structure the kernels rely on: no cross-component velocity coupling, B only between velocity dofs, twins present
    for (name, r, c) in keys:
        if r[2] < 2 and c[2] < 2:
            assert r[2] == c[2], "cross-component velocity entry"
        if name != "A":
            assert r[2] < 2 and c[2] < 2
    return keys


def has(keys, name, row, col):
    """Is M[row, col] in the union pattern?  `keys` lists the rows of ONE cell; by translation invariance of the pattern the
    row is shifted into the cell it belongs to (origin 2 * floor(d / 2))."""
    sx, sy = 2 * (row[0] // 2), 2 * (row[1] // 2)
    return (name, (row[0] - sx, row[1] - sy, row[2]), (col[0] - sx, col[1] - sy, col[2])) in keys


class Table:
    """Coefficient-table layout: entry i holds M[row, col] (times the branch sign when `signed`)."""

    def __init__(self):
        self.desc = []

    def add(self, mat, row, col, signed=False, twin=False, div=1, first=True):
        self.desc.append((mat, row, col, signed, twin, div, first))
        return len(self.desc) - 1


PARTS = (("A", (-2, -1)), ("B", (0,)), ("C", (1, 2)))  # window rows of the three body parts (see emit())


def window_nodes(rows):
    return [(dx, dy) for dy in rows for dx in range(-2, 3)]


def gen_forward(keys):
    """Row-owned.  Per source node m: xI = alpha[I_m], xJ = alpha[J_m] feed rows (I_t, J_t) of every target node t with
    the SAME coefficient (A's velocity block is diag(K, K), B1 = diag(Dx, Dx), B2 = diag(Dy, Dy)); pressure columns feed the
    A sums; the pressure row of the vertex takes both components."""
    T = Table()
    parts = {}
    for part, rows in PARTS:
      body = parts.setdefault(part, [])
      for (dx, dy) in window_nodes(rows):
        stmts, use = [], [False, False]
        for ti, (tx, ty) in enumerate(TARGETS):
            for mi, name in enumerate(MATS):
                if (name, (tx, ty, 0), (dx, dy, 0)) in keys or (name, (tx, ty, 1), (dx, dy, 1)) in keys:
                    i = T.add(name, (tx, ty, 0), (dx, dy, 0), twin=True)
                    stmts.append(f"FV({mi}, {ti}, {i})")
                    use = [True, True]
        for cc in (0, 1):
            if ("A", (0, 0, 2), (dx, dy, cc)) in keys:
                i = T.add("A", (0, 0, 2), (dx, dy, cc))
                stmts.append(f"FS{'IJ'[cc]}({i})")
                use[cc] = True
        if stmts:
            loads = [f"LDX(x{'IJ'[c]}, {dx}, {dy}, {c})" for c in (0, 1) if use[c]]
            body.append("  BEGIN " + " ".join(loads + stmts) + " END")
      for (dx, dy) in window_nodes(rows):
        if dx % 2 or dy % 2:
            continue
        stmts = []
        for ti, (tx, ty) in enumerate(TARGETS):
            for tc in (0, 1):
                if ("A", (tx, ty, tc), (dx, dy, 2)) in keys:
                    i = T.add("A", (tx, ty, tc), (dx, dy, 2))
                    stmts.append(f"FP({ti}, {tc}, {i})")
        if ("A", (0, 0, 2), (dx, dy, 2)) in keys:
            i = T.add("A", (0, 0, 2), (dx, dy, 2))
            stmts.append(f"FSP({i})")
        if stmts:
            body.append(f"  BEGIN LDX(xP, {dx}, {dy}, 2) " + " ".join(stmts) + " END")
    return T, parts


def element_multiplicity():
    """{(mat, row(tx,ty,comp), col(dx,dy,comp)): number of ELEMENTS that contribute a non-zero local entry} for the rows of one
    interior cell: what a matrix-free element walk (north_star: gather by dofmap, contract with the element tensor) evaluates
    separately and the assembled stencil has already summed.  From the fixture's own element integrals."""
    from feonet_navier_stokes_b200.fixtures.taylor_hood import _element_integrals

    n = 4
    mesh = structured_mesh(n)
    m = 2 * n + 1
    ox, oy = 4, 4  # an interior cell
    mult = {}
    tol = 1e-12

    def off(node):
        return (int(node % m) - ox, int(node // m) - oy)

    for sl, Kdd, D, Q, _ in _element_integrals(mesh):
        lap = Kdd[:, 0, 0] + Kdd[:, 1, 1]
        for e in range(lap.shape[0]):
            nodes = [off(v) for v in mesh.tri_p2[sl][e]]
            verts = [off(mesh.p1_to_p2[v]) for v in mesh.tri_p1[sl][e]]
            for a, ta in enumerate(nodes):
                if ta in TARGETS:
                    for b, mb in enumerate(nodes):
                        for name, val in (("A", lap[e, a, b]), ("B1", D[e, 0, a, b]), ("B2", D[e, 1, a, b])):
                            if abs(val) > tol:
                                for c in (0, 1):
                                    k = (name, (ta[0], ta[1], c), (mb[0], mb[1], c))
                                    mult[k] = mult.get(k, 0) + 1
                    for j, vq in enumerate(verts):
                        for c in (0, 1):
                            if abs(Q[e, c, a, j]) > tol:
                                k = ("A", (ta[0], ta[1], c), (vq[0], vq[1], 2))
                                mult[k] = mult.get(k, 0) + 1
            for j, vq in enumerate(verts):
                if vq == (0, 0):
                    for b, mb in enumerate(nodes):
                        for c in (0, 1):
                            if abs(Q[e, c, b, j]) > tol:
                                k = ("A", (0, 0, 2), (mb[0], mb[1], c))
                                mult[k] = mult.get(k, 0) + 1
    return mult


def gen_forward_element(keys):
    """The forward body of a matrix-free ELEMENT walk in gather form (row-owned, no scatter): every element incident to a target
    node contributes its own local row, i.e. one FMA per (element, local row, local column, matrix) instead of one per assembled
    entry.  Each assembled coefficient is split evenly over the element statements that sum to it (the planner only has the
    assembled CSR), so the results are the assembled ones up to round-off while the arithmetic COST is the element walk's.
    Used for the A/B measurement of DESIGN.md section 3.5 (FEO_LATTICE_ELEMENT=1), never by default."""
    mult = element_multiplicity()
    T = Table()
    parts = {}

    for part, rows in PARTS:
        body = parts.setdefault(part, [])
        for (dx, dy) in window_nodes(rows):
            stmts, use = [], [False, False]
            for ti, (tx, ty) in enumerate(TARGETS):
                for mi, name in enumerate(MATS):
                    present = (name, (tx, ty, 0), (dx, dy, 0)) in keys or (name, (tx, ty, 1), (dx, dy, 1)) in keys
                    k = mult.get((name, (tx, ty, 0), (dx, dy, 0)), 0)
                    if present or k:
                        k = max(k, 1)
                        for r in range(k):
                            i = T.add(name, (tx, ty, 0), (dx, dy, 0), twin=True, div=k, first=r == 0)
                            stmts.append(f"FV({mi}, {ti}, {i})")
                        use = [True, True]
            for cc in (0, 1):
                k = mult.get(("A", (0, 0, 2), (dx, dy, cc)), 0)
                if ("A", (0, 0, 2), (dx, dy, cc)) in keys or k:
                    k = max(k, 1)
                    for r in range(k):
                        i = T.add("A", (0, 0, 2), (dx, dy, cc), div=k, first=r == 0)
                        stmts.append(f"FS{'IJ'[cc]}({i})")
                    use[cc] = True
            if stmts:
                loads = [f"LDX(x{'IJ'[c]}, {dx}, {dy}, {c})" for c in (0, 1) if use[c]]
                body.append("  BEGIN " + " ".join(loads + stmts) + " END")
        for (dx, dy) in window_nodes(rows):
            if dx % 2 or dy % 2:
                continue
            stmts = []
            for ti, (tx, ty) in enumerate(TARGETS):
                for tc in (0, 1):
                    k = mult.get(("A", (tx, ty, tc), (dx, dy, 2)), 0)
                    if ("A", (tx, ty, tc), (dx, dy, 2)) in keys or k:
                        k = max(k, 1)
                        for r in range(k):
                            i = T.add("A", (tx, ty, tc), (dx, dy, 2), div=k, first=r == 0)
                            stmts.append(f"FP({ti}, {tc}, {i})")
            if ("A", (0, 0, 2), (dx, dy, 2)) in keys:
                i = T.add("A", (0, 0, 2), (dx, dy, 2))
                stmts.append(f"FSP({i})")
            if stmts:
                body.append(f"  BEGIN LDX(xP, {dx}, {dy}, 2) " + " ".join(stmts) + " END")
    return T, parts


def gen_backward(keys):
    """Column-owned.  Per source ROW node m: rI, rJ = r[I_m], r[J_m], d1, d2 = alpha[I_m], alpha[J_m]:
         T = a + b1 d1 + b2 d2 with a = A[I_m, I_t], b1 = s B1[I_m, I_t], b2 = s B2[I_m, I_t]:  g_I[t] += rI T, g_J[t] += rJ T
         Bu1_I[t] += f1 d1, Bu1_J[t] += f1 d2, Bu2_I[t] += f2 d1, Bu2_J[t] += f2 d2 with f1 = B1[I_t, I_m], f2 = B2[I_t, I_m]
         pressure column of the vertex:  g_P += A[I_m, P] rI + A[J_m, P] rJ
       per source pressure ROW q: rP = r[p_q]:  g_c[t] += A[p_q, c_t] rP,  g_P += A[p_q, P] rP."""
    T = Table()
    parts = {}
    for part, rows in PARTS:
      body = parts.setdefault(part, [])
      for (dx, dy) in window_nodes(rows):
        stmts, use_r, use_d = [], False, False
        for ti, (tx, ty) in enumerate(TARGETS):
            idx = []
            for name in MATS:
                present = has(keys, name, (dx, dy, 0), (tx, ty, 0)) or has(keys, name, (dx, dy, 1), (tx, ty, 1))
                idx.append(T.add(name, (dx, dy, 0), (tx, ty, 0), signed=name != "A", twin=True) if present else -1)
            if idx[1] >= 0 or idx[2] >= 0:
                stmts.append(f"BTB({ti}, {idx[0]}, {idx[1]}, {idx[2]})")
                use_r = use_d = True
            elif idx[0] >= 0:
                stmts.append(f"BTA({ti}, {idx[0]})")
                use_r = True
            for mi, name in ((1, "B1"), (2, "B2")):
                if (name, (tx, ty, 0), (dx, dy, 0)) in keys or (name, (tx, ty, 1), (dx, dy, 1)) in keys:
                    i = T.add(name, (tx, ty, 0), (dx, dy, 0), twin=True)
                    stmts.append(f"BF({mi}, {ti}, {i})")
                    use_d = True
        for cc in (0, 1):
            if has(keys, "A", (dx, dy, cc), (0, 0, 2)):
                i = T.add("A", (dx, dy, cc), (0, 0, 2))
                stmts.append(f"BS{'IJ'[cc]}({i})")
                use_r = True
        if stmts:
            loads = []
            if use_r:
                loads += [f"LDR(rI, {dx}, {dy}, 0)", f"LDR(rJ, {dx}, {dy}, 1)"]
            if use_d:
                loads += [f"LDA(d1, {dx}, {dy}, 0)", f"LDA(d2, {dx}, {dy}, 1)"]
            body.append("  BEGIN " + " ".join(loads + stmts) + " END")
      for (dx, dy) in window_nodes(rows):
        if dx % 2 or dy % 2:
            continue
        stmts = []
        for ti, (tx, ty) in enumerate(TARGETS):
            for tc in (0, 1):
                if has(keys, "A", (dx, dy, 2), (tx, ty, tc)):
                    i = T.add("A", (dx, dy, 2), (tx, ty, tc))
                    stmts.append(f"BP({ti}, {tc}, {i})")
        if has(keys, "A", (dx, dy, 2), (0, 0, 2)):
            i = T.add("A", (dx, dy, 2), (0, 0, 2))
            stmts.append(f"BSP({i})")
        if stmts:
            body.append(f"  BEGIN LDR(rP, {dx}, {dy}, 2) " + " ".join(stmts) + " END")
    return T, parts


def emit(keys):
    mat_id = {"A": 0, "B1": 1, "B2": 2}
    lines = [
        "// GENERATED by tools/gen_lattice_stencil.py -- do not edit.  Macro-stencil of the lattice plan (feo_lattice.h):",
        "// coefficient-table layouts and the straight-line cell bodies for the right-diagonal structured P2-P1 lattice.",
        "// DESC(index, matrix (0 A, 1 B1, 2 B2), row dx, dy, comp, col dx, dy, comp, signed, twin, div, first):",
        "//   table[index] = M[row, col] / div (times the branch sign when signed); comp 0 = u1 (I), 1 = u2 (J), 2 = pressure; node",
        "//   offsets are relative to the cell origin (2 ci, 2 cj); twin = the (J, J) entry of the same node pair must carry the same",
        "//   value; div > 1 only in the element-walk table FWDE, where `div` statements share one assembled entry (first marks one).",
        "// Bodies: one BEGIN ... END group per source node: LDX / LDR / LDA(var, dx, dy, comp) gather a line of alpha (forward),",
        "//   r and alpha (backward), then",
        "//   forward:  FV(matrix, target, i) FSI(i) FSJ(i) FP(target, comp, i) FSP(i)",
        "//   backward: BTA(target, ia) BTB(target, ia, ib1, ib2) BF(matrix, target, i) BSI(i) BSJ(i) BP(target, comp, i) BSP(i)",
        "//   (an index of -1 = entry absent);",
        "//   targets 0..3 = the cell's nodes V (0,0), H (1,0), T (0,1), D (1,1).",
        "// Each body comes in three parts by window row: A = rows -2, -1 (the rows a step releases first), B = row 0,",
        "//   C = rows 1, 2 (the rows a step receives last) -- the kernels release / wait for staged rows between the parts.",
    ]
    for tag, gen in (("FWD", gen_forward), ("BWD", gen_backward), ("FWDE", gen_forward_element)):
        T, body = gen(keys)
        lines.append(f"#define FEO_LAT_{tag}_NCOEF {len(T.desc)}")
        lines.append(f"#define FEO_LAT_{tag}_DESC(DESC) \\")
        for i, (mat, row, col, signed, twin, div, first) in enumerate(T.desc):
            lines.append(
                f"  DESC({i}, {mat_id[mat]}, {row[0]}, {row[1]}, {row[2]}, {col[0]}, {col[1]}, {col[2]}, {int(signed)}, {int(twin)}, {div}, {int(first)}) \\"
            )
        lines.append("")
        for part, _ in PARTS:
            lines.append(f"#define FEO_LAT_{tag}_BODY_{part} \\")
            for b in body[part]:
                lines.append(b + " \\")
            lines.append("")
    with open(OUT, "w") as f:
        f.write("\n".join(lines) + "\n")
    return lines


if __name__ == "__main__":
    keys = union_pattern()
    emit(keys)
    tf, bf = gen_forward(keys)
    tb, bb = gen_backward(keys)
    te, be = gen_forward_element(keys)
    print(f"union pattern: {len(keys)} matrix entries per cell; forward table {len(tf.desc)} coefficients, "
          f"backward table {len(tb.desc)} coefficients, element-walk forward table {len(te.desc)} coefficients -> {OUT}")
