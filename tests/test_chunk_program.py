"""Host-side model check of the chunk program of the tensor-core appearance kernel (csrc/appearance_mma_defs.cuh
build_program, exported as t2n_debug_chunk_program): the table the producer, issuer and loader roles walk, and the
mbarrier protocol the producers run on top of it (csrc/appearance_mma.cuh), for every decoder shape inside the
tensor-core envelope -- the GPU tests exercise only a handful of shapes.

Invariants checked by simulating the producers' bookkeeping (it, n0, done_known, s0_last) exactly as the kernel keeps it:
  * structure: all S2 chunks precede S1 chunk 0, which precedes Pro and every gather unit; S1 chunks and units appear in
    order; unit 2c+1 (which publishes basis chunk c) follows unit 2c;
  * every wait targets a chunk this warp has already published (no self-deadlock) and lies inside the 4-slot window of
    the chunk-done ring (p >= it - 4: the slot's phase cannot have been reused);
  * when a chunk is published, chunk it - 4 is known complete (the a_full slot is free);
  * the TMEM A stage (it % 3) and the shared-memory A stage (n0 & 1) written for a chunk were last used by a chunk that
    is known complete."""
import ctypes as C
import itertools

import pytest

from text2nerf_b200 import _native as nat

S2, S1, U, PRE, RAY, PRO, S3 = range(7)


def program(n_app_total, Kp):
    buf = (C.c_ubyte * 80)()
    n = nat.load().t2n_debug_chunk_program(n_app_total, Kp, buf, 80)
    assert n > 0, (n_app_total, Kp, n)
    return [(b & 7, b >> 3) for b in buf[:n]]


def shapes():
    for n_app_total in range(16, 161, 16):
        for pc, F in itertools.product((1, 2), range(0, 11)):
            if F == 0 and pc == 2:
                continue
            yield n_app_total, 32 * (1 + F * pc)


@pytest.mark.parametrize("n_app_total,Kp", list(shapes()))
def test_program_structure(n_app_total, Kp):
    prog = program(n_app_total, Kp)
    nk0, nk1, nk2 = (n_app_total + 31) // 32, Kp // 32, 4
    kinds = [k for k, _ in prog]
    assert kinds[0] == PRE and kinds[-1] == S3 and kinds.count(PRO) == 1 and kinds.count(RAY) == 1
    assert [i for k, i in prog if k == S2] == list(range(nk2))
    assert [i for k, i in prog if k == S1] == list(range(nk1))
    assert [i for k, i in prog if k == U] == list(range(2 * nk0))
    pos = {step: n for n, step in enumerate(prog)}
    assert max(pos[(S2, c)] for c in range(nk2)) < pos[(S1, 0)] < pos[(PRO, 0)] < pos[(U, 0)]
    assert pos[(RAY, 0)] < pos[(PRO, 0)] and pos[(PRE, 0)] < pos[(RAY, 0)]
    # the producers' straight-line head is Pre, S2 [0, early), Ray, S2 [early, nk2), S1 chunk 0, Pro
    assert prog[:nk2 + 4] == [(PRE, 0), (S2, 0), (S2, 1), (RAY, 0), (S2, 2), (S2, 3), (S1, 0), (PRO, 0)]
    assert len(prog) <= 80


def simulate(prog, n_tiles):
    it, n0, done_known, s0_last = 0, 0, -1, [-1, -1]
    published = []                      # (kind, tile, idx) in publish order
    last_ts_stage_user = {}

    def wait_done(p):
        nonlocal done_known
        if p > done_known:
            assert 0 <= p < it, "waits on a chunk this warp has not published"
            assert p >= it - 4, "chunk-done slot window exceeded"
            done_known = p

    for j in range(n_tiles + 2):
        ok = {S2: 2 <= j < n_tiles + 2, S1: 1 <= j < n_tiles + 1, U: j < n_tiles}
        for kind, idx in prog:
            if kind in (PRE, RAY, PRO, S3) or not ok[kind]:
                continue
            if kind in (S1, S2):
                wait_done(it - 3)
                stage = it % 3
                assert last_ts_stage_user.get(stage, -1) <= done_known, "TMEM A stage still in use"
                last_ts_stage_user[stage] = it
                assert it - 4 <= done_known, "a_full slot of chunk it-4 not free"
                published.append((kind, j - (2 if kind == S2 else 1), idx))
                it += 1
            else:
                st = n0 & 1
                if idx % 2 == 0:
                    wait_done(s0_last[st])
                    assert s0_last[st] <= done_known, "shared-memory A stage still in use"
                else:
                    wait_done(it - 4)
                    assert it - 4 <= done_known
                    s0_last[st] = it
                    n0 += 1
                    published.append((U, j, idx // 2))
                    it += 1
    return published


def roles_order(prog, n_tiles):
    """Chunk order as the issuer / loader derive it from the same table (is_chunk && step_ok)."""
    out = []
    for j in range(n_tiles + 2):
        for kind, idx in prog:
            if kind > U or (kind == U and idx % 2 == 0):
                continue
            t = j - (2 if kind == S2 else 1 if kind == S1 else 0)
            if 0 <= t < n_tiles:
                out.append((kind, t, idx // 2 if kind == U else idx))
    return out


@pytest.mark.parametrize("n_tiles", [1, 2, 3, 5])
def test_protocol_invariants_for_every_shape(n_tiles):
    for n_app_total, Kp in shapes():
        prog = program(n_app_total, Kp)
        published = simulate(prog, n_tiles)
        assert published == roles_order(prog, n_tiles), (n_app_total, Kp)
        nk0, nk1 = (n_app_total + 31) // 32, Kp // 32
        assert len(published) == n_tiles * (nk0 + nk1 + 4)


# ---- decoder-column recipe of the tensor-core path ----------------------------------------------------------------
import torch  # noqa: E402

from oracle import t2n_oracle as orc  # noqa: E402

HEADS = {"MLP_Fea_noview": 0, "MLP_Fea": 1, "MLP": 2}


def mma_recipe(mode, app_dim, fea_pe, view_pe):
    buf = (C.c_int * 1024)()
    n = nat.load().t2n_debug_mma_recipe(HEADS[mode], app_dim, fea_pe, view_pe, buf, 1024)
    if n < 0:
        return None
    v = list(buf[:n])
    return dict(n_freq=v[0], pe_chunks=v[1], Kp=v[2], ident_src=v[3:35], pe_src=v[35:67], pe_nf=v[67:99], perm=v[99:99 + v[2]])


@pytest.mark.parametrize("mode,app_dim,fea_pe,view_pe", [
    ("MLP_Fea_noview", 27, 6, 2), ("MLP_Fea_noview", 27, 0, 0), ("MLP_Fea_noview", 27, 10, 0), ("MLP_Fea_noview", 12, 3, 0),
    ("MLP_Fea", 27, 2, 2), ("MLP_Fea", 27, 6, 4), ("MLP_Fea", 13, 5, 1), ("MLP_Fea", 27, 3, 0), ("MLP_Fea", 27, 0, 3),
    ("MLP", 27, 6, 6), ("MLP", 27, 0, 4), ("MLP", 27, 0, 0), ("MLP", 29, 0, 10)])
def test_mma_recipe_produces_every_reference_column_once(mode, app_dim, fea_pe, view_pe):
    """Columns generated the way app_forward_mma_kernel generates them (chunk 0 = identity columns, chunk 1 + f*pc + h =
    (sin, cos)(base[pe_src[16h + e]] * 2^f)) and sent through `perm` reproduce the reference's decoder input
    (tensorBase.py:62-109, 137-159 via oracle.freq_encode) column for column; everything else has zero weight."""
    R = mma_recipe(mode, app_dim, fea_pe, view_pe)
    assert R is not None
    A = app_dim
    g = torch.Generator().manual_seed(1)
    feat, view, xn = torch.randn(4, A, generator=g), torch.randn(4, 3, generator=g), torch.randn(4, 3, generator=g)
    base = torch.cat([feat, view, xn, torch.zeros(4, 1)], -1).double()
    assert R["Kp"] == 32 * (1 + R["n_freq"] * R["pe_chunks"]) and len(R["perm"]) == R["Kp"]
    internal = torch.zeros(4, R["Kp"], dtype=torch.float64)
    for k in range(32):
        internal[:, k] = base[:, R["ident_src"][k]]
    valid = torch.zeros(R["Kp"], dtype=torch.bool)
    valid[:32] = True
    for f in range(R["n_freq"]):
        for h in range(R["pe_chunks"]):
            for e1 in range(16):
                e = 16 * h + e1
                k = 32 * (1 + f * R["pe_chunks"] + h) + 2 * e1
                ang = base[:, R["pe_src"][e]] * float(1 << f)
                internal[:, k], internal[:, k + 1] = torch.sin(ang), torch.cos(ang)
                valid[k] = valid[k + 1] = f < R["pe_nf"][e]
    ref_cols = [feat] + ([view] if mode != "MLP_Fea_noview" else [])
    if mode in ("MLP_Fea_noview", "MLP_Fea") and fea_pe > 0:
        ref_cols.append(orc.freq_encode(feat.double(), fea_pe))
    if mode in ("MLP_Fea", "MLP") and view_pe > 0:
        ref_cols.append(orc.freq_encode(view.double(), view_pe))
    ref = torch.cat([c.double() for c in ref_cols], -1)
    used = [p for p in R["perm"] if p >= 0]
    assert sorted(used) == list(range(ref.shape[1])), "every reference column exactly once"
    for k, src in enumerate(R["perm"]):
        if src >= 0:
            assert bool(valid[k]), (k, src)
            assert torch.allclose(internal[:, k], ref[:, src], rtol=0, atol=1e-12), (k, src)


def test_mma_recipe_rejects_heads_outside_the_envelope():
    assert mma_recipe("MLP_Fea_noview", 30, 6, 0) is None        # app_dim > 29
    assert mma_recipe("MLP_Fea", 27, 11, 0) is None              # more than 10 frequencies
    assert nat.load().t2n_debug_mma_recipe(3, 27, 0, 0, (C.c_int * 1024)(), 1024) < 0      # SH head: no decoder


@pytest.mark.parametrize("mode,app_dim,fea_pe,view_pe", [
    ("MLP_Fea_noview", 27, 6, 2), ("MLP_Fea_noview", 27, 0, 0), ("MLP_Fea_noview", 12, 3, 0), ("MLP_Fea", 27, 2, 2),
    ("MLP_Fea", 27, 6, 4), ("MLP_Fea", 13, 5, 1), ("MLP_Fea", 27, 0, 3), ("MLP", 27, 6, 6), ("MLP", 27, 0, 0), ("MLP", 29, 0, 10)])
def test_backward_recipe_is_a_reordering_of_the_forward_identity_chunk(mode, app_dim, fea_pe, view_pe):
    """The backward kernels keep the forward's PE chunks and re-order the identity chunk so that the thread owning PE
    entries {4ph.., 16+4ph..} also owns the identity columns of the same base entries (bwd_mma_defs.cuh): every reference
    column must still be covered exactly once, identity slot s must point at the base entry of its PE entry, and no base
    entry may appear in two slots."""
    R = mma_recipe(mode, app_dim, fea_pe, view_pe)
    buf = (C.c_int * 1024)()
    n = nat.load().t2n_debug_mma_bwd_recipe(HEADS[mode], app_dim, fea_pe, view_pe, buf, 1024)
    assert n == 32 + R["Kp"]
    own, perm = list(buf[:32]), list(buf[32:n])
    zero = app_dim + 6
    assert perm[32:] == R["perm"][32:]                                  # PE chunks unchanged
    assert sorted(p for p in perm if p >= 0) == sorted(p for p in R["perm"] if p >= 0)
    for s in range(32):
        ph, q = s >> 3, s & 7
        e = 4 * ph + q if q < 4 else 16 + 4 * ph + (q - 4)
        if R["pe_nf"][e] > 0:
            assert own[s] == R["pe_src"][e], (s, e)
    real = [b for b in own if b != zero]
    assert len(real) == len(set(real)), "a base entry owned by two slots"
    fwd_ident = {R["ident_src"][k]: R["perm"][k] for k in range(32) if R["perm"][k] >= 0}
    for s in range(32):
        assert perm[s] == fwd_ident.get(own[s], -1), s
    assert set(fwd_ident) <= set(real)
