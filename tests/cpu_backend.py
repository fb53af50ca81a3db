"""numpy emulation of the libpolydis_b200 C-ABI on HOST pointers -- test infrastructure only.

The product has no CPU path.  To test the host-side logic (autograd orchestration, teacher-forcing
plans, layouts, strides) in the GPU-less build container, the ``-m "not gpu"`` tests swap
``polydis_b200._lib.call`` for ``CpuBackend.call``: each entry point is re-stated here in numpy on the
raw pointers + strides the host code passes, exactly per the contract in include/polydis_b200.h.
The ``-m gpu`` tests then check each CUDA kernel against these same emulations and the whole model
against the oracle.
"""
import ctypes

import numpy as np


def _arr(ptr, shape, strides, dtype=np.float32):
    if ptr is None:
        return None
    shape = tuple(int(s) for s in shape)
    strides = tuple(int(s) for s in strides)
    item = np.dtype(dtype).itemsize
    if any(s == 0 for s in shape):
        return np.zeros(shape, dtype=dtype)
    span = sum((s - 1) * st for s, st in zip(shape, strides)) + 1
    buf = (ctypes.c_char * (span * item)).from_address(int(ptr))
    flat = np.frombuffer(buf, dtype=dtype)
    return np.lib.stride_tricks.as_strided(flat, shape, tuple(st * item for st in strides))


def _sig(x):
    return 1.0 / (1.0 + np.exp(-x, dtype=np.float32))


class CpuBackend:
    def __init__(self):
        self.calls = []

    def call(self, name, *a):
        self.calls.append(name)
        getattr(self, name)(*a)

    # ---- gemm ----
    def pd_gemm_f32(self, A, sam, sak, B, sbk, sbn, C, ldc, bias, M, N, K, acc, st):
        if M <= 0 or N <= 0:
            return
        a = _arr(A, (M, K), (sam, sak))
        b = _arr(B, (K, N), (sbk, sbn))
        c = _arr(C, (M, N), (ldc, 1))
        r = (a @ b).astype(np.float32) if K > 0 else np.zeros((M, N), np.float32)
        if bias is not None:
            r = r + _arr(bias, (N,), (1,))
        c[...] = c + r if acc else r

    def pd_gemm_tf32(self, *a):
        # same contract; the emulation keeps fp32 products (the GPU kernel rounds operands to TF32)
        self.pd_gemm_f32(*a)

    def pd_gemm_tf32_cfg(self, A, sam, sak, B, sbk, sbn, C, ldc, bias, M, N, K, acc, cfg, st):
        self.pd_gemm_f32(A, sam, sak, B, sbk, sbn, C, ldc, bias, M, N, K, acc, st)

    @staticmethod
    def _bf16_to_f32(ptr, shape, strides):
        u = _arr(ptr, shape, strides, np.uint16).astype(np.uint32) << 16
        return u.view(np.float32)

    def pd_gemm_bf16(self, A, sam, sak, B, sbk, sbn, C, ldc, bias, M, N, K, acc, st):
        if M <= 0 or N <= 0:
            return
        a = self._bf16_to_f32(A, (M, K), (sam, sak))
        b = self._bf16_to_f32(B, (K, N), (sbk, sbn))
        c = _arr(C, (M, N), (ldc, 1))
        r = (a @ b).astype(np.float32)
        if bias is not None:
            r = r + _arr(bias, (N,), (1,))
        c[...] = c + r if acc else r

    def pd_f32_to_bf16(self, x, ldx, rows, cols, out, ldo, st):
        X = _arr(x, (rows, cols), (ldx, 1)).astype(np.float32)
        bits = X.view(np.uint32).astype(np.uint64)
        rounded = (bits + 0x7FFF + ((bits >> 16) & 1)) >> 16                     # round to nearest even
        _arr(out, (rows, cols), (ldo, 1), np.uint16)[...] = rounded.astype(np.uint16)

    def pd_tf32_split(self, x, ldx, rows, cols, hi_out, lo, ldo, st):
        X = _arr(x, (rows, cols), (ldx, 1)).astype(np.float32)
        bits = X.view(np.uint32).astype(np.uint64)
        hi = (((bits + 0x1000) >> 13) << 13).astype(np.uint32).view(np.float32)     # round-to-nearest (ties away)
        _arr(hi_out, (rows, cols), (ldo, 1))[...] = hi
        _arr(lo, (rows, cols), (ldo, 1))[...] = X - hi

    def pd_tf32_split3(self, x, ldx, rows, cols, out, ldo, order, st):
        kp = (cols + 3) // 4 * 4
        X = np.zeros((rows, kp), np.float32)
        X[:, :cols] = _arr(x, (rows, cols), (ldx, 1))
        bits = X.view(np.uint32).astype(np.uint64)
        hi = (((bits + 0x1000) >> 13) << 13).astype(np.uint32).view(np.float32)
        O = _arr(out, (rows, 3 * kp), (ldo, 1))
        O[:, :kp] = hi
        O[:, kp:2 * kp] = (X - hi) if order else hi
        O[:, 2 * kp:] = hi if order else (X - hi)

    def pd_colsum_f32(self, X, ldx, M, N, out, acc, st):
        o = _arr(out, (N,), (1,))
        s = _arr(X, (M, N), (ldx, 1)).sum(0, dtype=np.float32) if M > 0 else 0.0
        o[...] = o + s if acc else s

    def pd_colsum_seq_f32(self, X, ldx, R, T, N, lengths, out, acc, st):
        o = _arr(out, (N,), (1,))
        x = _arr(X, (R, T, N), (T * ldx, ldx, 1))
        live = np.arange(T)[None, :] < _arr(lengths, (R,), (1,), np.int32)[:, None]
        s = np.where(live[:, :, None], x, 0.0).sum((0, 1), dtype=np.float32)       # dead rows are never read (may hold NaN)
        o[...] = o + s if acc else s

    def pd_sum_steps_f32(self, X, ldr, ldt, T, out, ldo, R, C, st):
        _arr(out, (R, C), (ldo, 1))[...] = _arr(X, (R, T, C), (ldr, ldt, 1)).sum(1, dtype=np.float32)

    def pd_transpose_f32(self, inp, rows, cols, out, st):
        _arr(out, (cols, rows), (rows, 1))[...] = _arr(inp, (rows, cols), (cols, 1)).T

    # ---- gru gates ----
    def pd_gru_gates_fwd_split3(self, gi, ldgi, gi2, ldgi2, gh, ldgh, hp, ldhp, ho, ldho, lengths, t, B, H, h3, ldh3, st):
        self.pd_gru_gates_fwd(gi, ldgi, gi2, ldgi2, gh, ldgh, hp, ldhp, ho, ldho, None, 0, None, 0, lengths, t, B, H, st)
        self.pd_tf32_split3(ho, ldho, B, H, h3, ldh3, 0, st)

    def pd_gru_gates_fwd(self, gi, ldgi, gi2, ldgi2, gh, ldgh, hp, ldhp, ho, ldho, rzn, ldrzn, hn, ldhn,
                         lengths, t, B, H, st):
        GI = _arr(gi, (B, 3 * H), (ldgi, 1)).copy()
        if gi2 is not None:
            GI = GI + _arr(gi2, (B, 3 * H), (ldgi2, 1))
        GH = _arr(gh, (B, 3 * H), (ldgh, 1))
        HP = _arr(hp, (B, H), (ldhp, 1)) if hp is not None else np.zeros((B, H), np.float32)
        r = _sig(GI[:, :H] + GH[:, :H])
        z = _sig(GI[:, H:2 * H] + GH[:, H:2 * H])
        n = np.tanh(GI[:, 2 * H:] + r * GH[:, 2 * H:])
        h = (1 - z) * n + z * HP
        act = np.ones(B, bool) if lengths is None else (t < _arr(lengths, (B,), (1,), np.int32))
        HO = _arr(ho, (B, H), (ldho, 1))
        HO[...] = np.where(act[:, None], h, HP)
        if rzn is not None:
            S = _arr(rzn, (B, 3 * H), (ldrzn, 1))
            S[act] = np.concatenate([r, z, n], 1)[act]
        if hn is not None:
            Q = _arr(hn, (B, H), (ldhn, 1))
            Q[act] = GH[:, 2 * H:][act]

    def pd_gru_step_tma(self, hp, ldhp, w, ldw, b_hh, gi, ldgi, gi2, ldgi2, ho, ldho, rzn, ldrzn, hn, ldhn, B, H, st):
        self.pd_gru_step_tf32(hp, ldhp, w, ldw, b_hh, gi, ldgi, gi2, ldgi2, ho, ldho, rzn, ldrzn, hn, ldhn, None, 0, B, H, st)

    def pd_gru_step_tma3(self, a3, lda3, w3, ldw3, b_hh, gi, ldgi, gi2, ldgi2, hp, ldhp, ho, ldho, h3, ldh3, B, H, st):
        A = _arr(a3, (B, 3 * H), (lda3, 1)).astype(np.float64)
        W = _arr(w3, (3 * H, 3 * H), (ldw3, 1)).astype(np.float64)
        gh = (A @ W.T + _arr(b_hh, (3 * H,), (1,))).astype(np.float32)     # hi.hi + hi.lo + lo.hi, fp32-class
        self.pd_gru_gates_fwd(gi, ldgi, gi2, ldgi2, gh.ctypes.data, 3 * H, hp, ldhp, ho, ldho, None, 0, None, 0, None, 0,
                              B, H, st)
        self.pd_tf32_split3(ho, ldho, B, H, h3, ldh3, 0, st)

    def pd_gru_step_tmax(self, hp, ldhp, w, ldw, x, ldx, wx, ldwx, K2, b_hh, gi2, ldgi2, ho, ldho, rzn, ldrzn, hn, ldhn, B, H, st):
        gi = (_arr(x, (B, K2), (ldx, 1)) @ _arr(wx, (3 * H, K2), (ldwx, 1)).T).astype(np.float32)
        self.pd_gru_step_tf32(hp, ldhp, w, ldw, b_hh, gi.ctypes.data, 3 * H, gi2, ldgi2, ho, ldho, rzn, ldrzn, hn, ldhn, None, 0,
                              B, H, st)

    def pd_gru_step_tma3x(self, a3, lda3, w3, ldw3, x3, ldx3, wx3, ldwx3, K2, b_hh, gi2, ldgi2, hp, ldhp, ho, ldho, h3, ldh3,
                          B, H, st):
        gi = (_arr(x3, (B, K2), (ldx3, 1)).astype(np.float64) @ _arr(wx3, (3 * H, K2), (ldwx3, 1)).astype(np.float64).T
              ).astype(np.float32)
        self.pd_gru_step_tma3(a3, lda3, w3, ldw3, b_hh, gi.ctypes.data, 3 * H, gi2, ldgi2, hp, ldhp, ho, ldho, h3, ldh3, B, H, st)

    def pd_gru_step_tf32(self, hp, ldhp, w, ldw, b_hh, gi, ldgi, gi2, ldgi2, ho, ldho, rzn, ldrzn, hn, ldhn, lengths, t,
                         B, H, st):
        HP = _arr(hp, (B, H), (ldhp, 1))
        gh = (HP @ _arr(w, (3 * H, H), (ldw, 1)).T + _arr(b_hh, (3 * H,), (1,))).astype(np.float32)
        gh_ptr = gh.ctypes.data
        self.pd_gru_gates_fwd(gi, ldgi, gi2, ldgi2, gh_ptr, 3 * H, hp, ldhp, ho, ldho, rzn, ldrzn, hn, ldhn, lengths, t,
                              B, H, st)

    def pd_gru128_fwd_perm(self, gi, ldr, ldt, lengths, w_hh, b_hh, h_all, hr, ht, rzn, zr, zt, hn, nr, nt, R, T, reverse,
                           passes, perm, st):
        p = _arr(perm, (R,), (1,), np.int32)
        assert sorted(p.tolist()) == list(range(R))          # (a visiting order: same arithmetic per row)
        self.pd_gru128_fwd(gi, ldr, ldt, lengths, w_hh, b_hh, h_all, hr, ht, rzn, zr, zt, hn, nr, nt, R, T, reverse, passes, st)

    def pd_gru128_fwd(self, gi, ldr, ldt, lengths, w_hh, b_hh, h_all, hr, ht, rzn, zr, zt, hn, nr, nt, R, T, reverse,
                      passes, st):
        H = 128
        GI = _arr(gi, (R, T, 3 * H), (ldr, ldt, 1))
        L = np.minimum(_arr(lengths, (R,), (1,), np.int32), T)
        W, b = _arr(w_hh, (3 * H, H), (H, 1)), _arr(b_hh, (3 * H,), (1,))
        HA = _arr(h_all, (R, T, H), (hr, ht, 1))
        Z = _arr(rzn, (R, T, 3 * H), (zr, zt, 1)) if rzn is not None else None
        N = _arr(hn, (R, T, H), (nr, nt, 1)) if hn is not None else None
        h = np.zeros((R, H), np.float32)
        for s_ in range(T):
            t = T - 1 - s_ if reverse else s_
            gh = (h @ W.T + b).astype(np.float32)
            g = GI[:, t]
            r = _sig(g[:, :H] + gh[:, :H])
            z = _sig(g[:, H:2 * H] + gh[:, H:2 * H])
            n = np.tanh(g[:, 2 * H:] + r * gh[:, 2 * H:])
            act = t < L
            h = np.where(act[:, None], (1 - z) * n + z * h, h).astype(np.float32)
            HA[:, t] = h
            if Z is not None:
                Z[act, t] = np.concatenate([r, z, n], 1)[act]
            if N is not None:
                N[act, t] = gh[:, 2 * H:][act]

    def pd_gru128_bwd(self, dout, dr, dt, h_all, hr, ht, rzn, zr, zt, hn, nr, nt, lengths, w_hh, dgi, gr, gt, dgh, qr,
                      qt, R, T, reverse, st):
        H = 128
        DO = _arr(dout, (R, T, H), (dr, dt, 1))
        HA = _arr(h_all, (R, T, H), (hr, ht, 1))
        Z = _arr(rzn, (R, T, 3 * H), (zr, zt, 1))
        N = _arr(hn, (R, T, H), (nr, nt, 1))
        L = np.minimum(_arr(lengths, (R,), (1,), np.int32), T)
        W = _arr(w_hh, (3 * H, H), (H, 1))
        GI = _arr(dgi, (R, T, 3 * H), (gr, gt, 1))
        GH = _arr(dgh, (R, T, 3 * H), (qr, qt, 1))
        dh = np.zeros((R, H), np.float32)
        for s_ in range(T - 1, -1, -1):
            t = T - 1 - s_ if reverse else s_
            tp = t + 1 if reverse else t - 1
            d = dh + DO[:, t]
            act = (t < L)[:, None]
            hp = HA[:, tp] if 0 <= tp < T else np.zeros((R, H), np.float32)
            with np.errstate(all="ignore"):
                r, z, n = Z[:, t, :H], Z[:, t, H:2 * H], Z[:, t, 2 * H:]
                dn = d * (1 - z) * (1 - n * n)
                dz = d * (hp - n) * z * (1 - z)
                drr = dn * N[:, t] * r * (1 - r)
                g_i = np.where(act, np.concatenate([drr, dz, dn], 1), 0).astype(np.float32)
                g_h = np.where(act, np.concatenate([drr, dz, dn * r], 1), 0).astype(np.float32)
            GI[:, t] = g_i
            GH[:, t] = g_h
            dh = (np.where(act, d * z, d) + g_h @ W).astype(np.float32)

    def pd_gru_gates_bwd(self, dh, lddh, dh2, lddh2, dh3, lddh3, rzn, ldrzn, hn, ldhn, hp, ldhp, dgi, lddgi, dgh, lddgh,
                         dhp, lddhp, dgi2, lddgi2, lengths, t, B, H, st):
        d = np.zeros((B, H), np.float32)
        if dh is not None:
            d = d + _arr(dh, (B, H), (lddh, 1))
        if dh2 is not None:
            d = d + _arr(dh2, (B, H), (lddh2, 1))
        if dh3 is not None:
            d = d + _arr(dh3, (B, H), (lddh3, 1))
        act = np.ones(B, bool) if lengths is None else (t < _arr(lengths, (B,), (1,), np.int32))
        S = _arr(rzn, (B, 3 * H), (ldrzn, 1))
        r, z, n = S[:, :H], S[:, H:2 * H], S[:, 2 * H:]
        HN = _arr(hn, (B, H), (ldhn, 1))
        HP = _arr(hp, (B, H), (ldhp, 1)) if hp is not None else np.zeros((B, H), np.float32)
        with np.errstate(all="ignore"):
            dn = d * (1 - z) * (1 - n * n)
            dz = d * (HP - n) * z * (1 - z)
            dr = dn * HN * r * (1 - r)
            dnr = dn * r
        a = act[:, None]
        G = _arr(dgi, (B, 3 * H), (lddgi, 1))
        G[...] = np.where(a, np.concatenate([dr, dz, dn], 1), 0)
        Gh = _arr(dgh, (B, 3 * H), (lddgh, 1))
        Gh[...] = np.where(a, np.concatenate([dr, dz, dnr], 1), 0)
        _arr(dhp, (B, H), (lddhp, 1))[...] = np.where(a, d * z, d)
        if dgi2 is not None:
            G2 = _arr(dgi2, (B, 3 * H), (lddgi2, 1))
            G2[...] = G2 + G

    # ---- packed note level (csrc/packed.cu and the *_rows entry points) ----
    # The emulations write exactly the rows the kernels write and leave the others untouched; the packed-path tests fill
    # every ``torch.empty`` buffer with NaN (``poison_empty``), so host logic that lets a dead row leak into a result fails.
    @staticmethod
    def _live(cp, slot_rows, n_rows, gran):
        """-> (processed, live) bool masks over n_rows slot-major rows: a row is processed iff its ``gran``-row block holds
        a live row (r % slot_rows < cp[r // slot_rows])."""
        n_slots = (n_rows + slot_rows - 1) // slot_rows
        CP = _arr(cp, (n_slots,), (1,), np.int32)
        q = np.arange(n_rows)
        live = (q % slot_rows) < CP[q // slot_rows]
        blk = np.zeros((n_rows + gran - 1) // gran, bool)
        np.logical_or.at(blk, q // gran, live)
        return blk[q // gran], live

    def pd_pack_order(self, lengths, R, perm, inv, table, st):
        L = np.clip(_arr(lengths, (R,), (1,), np.int32), 0, 16)
        order = np.argsort(-L, kind="stable").astype(np.int32)
        _arr(perm, (R,), (1,), np.int32)[...] = order
        iv = np.empty(R, np.int32)
        iv[order] = np.arange(R, dtype=np.int32)
        _arr(inv, (R,), (1,), np.int32)[...] = iv
        T = _arr(table, (64,), (1,), np.int32)
        for t in range(17):
            c = int((L > t).sum())
            cp = min(R, (c + 127) // 128 * 128)
            T[t], T[17 + t], T[34 + t] = c, cp, 6 * cp

    def pd_pack_grid(self, tok, lengths, perm, R, tok_s, pt_s, dt_s, len_s, st):
        T = _arr(tok, (R, 16, 6), (96, 6, 1), np.int32)
        P = _arr(perm, (R,), (1,), np.int32)
        S = T[P].transpose(1, 0, 2)                          # (16, R, 6)
        _arr(tok_s, (16, R, 6), (R * 6, 6, 1), np.int32)[...] = S
        _arr(pt_s, (15, R), (R, 1), np.int32)[...] = S[1:, :, 0]
        _arr(dt_s, (15, R, 5), (R * 5, 5, 1), np.int32)[...] = S[1:, :, 1:]
        _arr(len_s, (R,), (1,), np.int32)[...] = _arr(lengths, (R,), (1,), np.int32)[P]

    def pd_gather_rows_f32(self, src, lds, idx, R, C, dst, ldd, st):
        I = _arr(idx, (R,), (1,), np.int32)
        _arr(dst, (R, C), (ldd, 1))[...] = _arr(src, (R, C), (lds, 1))[I]

    def pd_sum_slots_rows_f32(self, X, ldt, ldr, T, cp, out, ldo, R, C, st):
        x = _arr(X, (T, R, C), (ldt, ldr, 1))
        CP = _arr(cp, (T,), (1,), np.int32)
        m = (np.arange(R)[None, :] < CP[:, None])[:, :, None]
        _arr(out, (R, C), (ldo, 1))[...] = np.where(m, x, 0).sum(0, dtype=np.float32)

    def pd_colsum_rows_f32(self, X, ldx, M, N, out, acc, cp, slot_rows, st):
        o = _arr(out, (N,), (1,))
        proc, _ = self._live(cp, slot_rows, M, 32)
        s = _arr(X, (M, N), (ldx, 1))[proc].sum(0, dtype=np.float32)
        o[...] = o + s if acc else s

    def pd_gemm_tf32_rows(self, A, sam, sak, B, sbk, sbn, C, ldc, bias, M, N, K, acc, pred, cp, slot_rows, st):
        a = _arr(A, (M, K), (sam, sak))
        b = _arr(B, (K, N), (sbk, sbn))
        c = _arr(C, (M, N), (ldc, 1))
        with np.errstate(all="ignore"):
            if pred == 1:
                proc, _ = self._live(cp, slot_rows, M, 128)
                r = (a[proc] @ b).astype(np.float32)
                if bias is not None:
                    r = r + _arr(bias, (N,), (1,))
                c[proc] = c[proc] + r if acc else r
            else:
                proc, _ = self._live(cp, slot_rows, K, 32)
                r = (a[:, proc] @ b[proc]).astype(np.float32)
                c[...] = c + r if acc else r

    def pd_gru_step_tmax_rows(self, hp, ldhp, w, ldw, x, ldx, wx, ldwx, K2, b_hh, gi2, ldgi2, ho, ldho, rzn, ldrzn, hn, ldhn,
                              B, H, nrows, st):
        live = min(B, int(_arr(nrows, (1,), (1,), np.int32)[0]))
        P = min(B, (live + 127) // 128 * 128)
        with np.errstate(all="ignore"):
            if P > 0:
                self.pd_gru_step_tmax(hp, ldhp, w, ldw, x, ldx, wx, ldwx, K2, b_hh, gi2, ldgi2, ho, ldho, rzn, ldrzn, hn, ldhn, P, H, st)

    def pd_gru_gates_bwd_rows(self, dh, lddh, dh2, lddh2, dh3, lddh3, rzn, ldrzn, hn, ldhn, hp, ldhp, dgi, lddgi, dgh, lddgh,
                              dhp, lddhp, B, H, nrows, st):
        P = min(B, int(_arr(nrows, (1,), (1,), np.int32)[0]))
        if P > 0:
            self.pd_gru_gates_bwd(dh, lddh, dh2, lddh2, dh3, lddh3, rzn, ldrzn, hn, ldhn, hp, ldhp, dgi, lddgi, dgh, lddgh,
                                  dhp, lddhp, None, 0, None, 0, P, H, st)

    def pd_dur_decode_fwd_rows(self, h0, ldh0, Q, w_ih, b_ih, w_hh, b_hh, sos, w_out, b_out, logits, S, cp, slot_rows, st):
        proc, _ = self._live(cp, slot_rows, Q, 16)
        L = _arr(logits, (Q, 5, 2), (10, 2, 1))
        Sb = _arr(S, (Q, 6, 72), (432, 72, 1)) if S is not None else None
        keep = (L[~proc].copy(), None if Sb is None else Sb[~proc].copy())
        with np.errstate(all="ignore"):
            self.pd_dur_decode_fwd(h0, ldh0, Q, w_ih, b_ih, w_hh, b_hh, sos, w_out, b_out, logits, S, 1, st)
        L[~proc] = keep[0]                                   # dead 16-note tiles are not written
        if Sb is not None:
            Sb[~proc] = keep[1]

    def pd_dur_decode_bwd_rows(self, S, dlog, Q, w_ih, b_ih, w_hh, b_hh, sos, w_out, b_out, GX, dh0, lddh0, cp, slot_rows, st):
        proc, _ = self._live(cp, slot_rows, Q, 16)
        G, D = _arr(GX, (Q, 6, 264), (6 * 264, 264, 1)), _arr(dh0, (Q, 64), (lddh0, 1))
        keep = (G[~proc].copy(), D[~proc].copy())
        with np.errstate(all="ignore"):
            self.pd_dur_decode_bwd(S, dlog, Q, w_ih, b_ih, w_hh, b_hh, sos, w_out, b_out, GX, dh0, lddh0, 1, st)
        G[~proc], D[~proc] = keep

    def pd_note_embed_bwd_rows(self, tok, R, g, ldg, dWT, db, cp, slot_rows, st):
        _, live = self._live(cp, slot_rows, R, 1)
        T = _arr(tok, (R, 6), (6, 1), np.int32)[live]
        G = _arr(g, (R, 128), (ldg, 1))[live]
        W = _arr(dWT, (135, 128), (128, 1))
        p = T[:, 0]
        ok = (p >= 0) & (p < 130)
        np.add.at(W, p[ok], G[ok])
        W[130:] += T[:, 1:].astype(np.float32).T @ G
        _arr(db, (128,), (1,))[...] += G.sum(0)

    def pd_gru128_bwd_rows(self, dout, dr, dt, h_all, hr, ht, rzn, zr, zt, hn, nr, nt, lengths, w_hh, dgi, gr, gt, dgh, qr,
                           qt, R, T, reverse, cp, dout_step, st):
        GI = _arr(dgi, (R, T, 384), (gr, gt, 1))
        keep = GI.copy()
        if dout_step >= 0:                                   # (R,128) gradient of one step's output
            full = np.zeros((R, T, 128), np.float32)
            full[:, dout_step] = _arr(dout, (R, 128), (dr, 1))
            dout, dr, dt = full.ctypes.data, T * 128, 128
        self.pd_gru128_bwd(dout, dr, dt, h_all, hr, ht, rzn, zr, zt, hn, nr, nt, lengths, w_hh, dgi, gr, gt, dgh, qr, qt, R, T,
                           reverse, st)
        if cp is None:
            return
        L = np.minimum(_arr(lengths, (R,), (1,), np.int32), T)
        CP = _arr(cp, (T,), (1,), np.int32)
        skip = (np.arange(T)[None, :] >= L[:, None]) & (np.arange(R)[:, None] >= CP[None, :])      # masked and not padded-live
        GI[skip] = keep[skip]

    @staticmethod
    def _to_bf16_bits(x):
        bits = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
        return ((bits + 0x7FFF + ((bits >> 16) & 1)) >> 16).astype(np.uint16)

    def pd_gru_step_tma_bf16(self, hb, ldhb, wb, ldwb, b_hh, gi, ldgi, gi2, ldgi2, hp, ldhp, ho, ldho, hbo, ldhbo, rzn, ldrzn,
                             hn, ldhn, B, H, st):
        HB = self._bf16_to_f32(hb, (B, H), (ldhb, 1))
        WB = self._bf16_to_f32(wb, (3 * H, H), (ldwb, 1))
        gh = (HB @ WB.T + _arr(b_hh, (3 * H,), (1,))).astype(np.float32)
        self.pd_gru_gates_fwd(gi, ldgi, gi2, ldgi2, gh.ctypes.data, 3 * H, hp, ldhp, ho, ldho, rzn, ldrzn, hn, ldhn, None, 0,
                              B, H, st)
        _arr(hbo, (B, H), (ldhbo, 1), np.uint16)[...] = self._to_bf16_bits(_arr(ho, (B, H), (ldho, 1)))

    def pd_gru_step_tma_bf16_units(self, hb, ldhb, wb, ldwb, b_hh, gi, ldgi, gi2, ldgi2, hp, ldhp, ho, ldho, hbo, ldhbo, rzn,
                                   ldrzn, hn, ldhn, B, H, units, st):
        assert units in (32, 64)                 # (a launch-shape choice: same arithmetic)
        self.pd_gru_step_tma_bf16(hb, ldhb, wb, ldwb, b_hh, gi, ldgi, gi2, ldgi2, hp, ldhp, ho, ldho, hbo, ldhbo, rzn, ldrzn,
                                  hn, ldhn, B, H, st)

    def pd_gru_gates_bwd_zb(self, dh, lddh, dh2, lddh2, dh3, lddh3, rzn, ldrzn, hn, ldhn, hp, ldhp, dgi, lddgi, dgh, lddgh,
                            dhp, lddhp, lengths, t, B, H, zero_out, ldzo, dgh_b, lddghb, st):
        self.pd_gru_gates_bwd_z(dh, lddh, dh2, lddh2, dh3, lddh3, rzn, ldrzn, hn, ldhn, hp, ldhp, dgi, lddgi, dgh, lddgh,
                                dhp, lddhp, lengths, t, B, H, zero_out, ldzo, st)
        _arr(dgh_b, (B, 3 * H), (lddghb, 1), np.uint16)[...] = self._to_bf16_bits(_arr(dgh, (B, 3 * H), (lddgh, 1)))

    def pd_gru_gates_bwd_z(self, dh, lddh, dh2, lddh2, dh3, lddh3, rzn, ldrzn, hn, ldhn, hp, ldhp, dgi, lddgi, dgh, lddgh,
                           dhp, lddhp, lengths, t, B, H, zero_out, ldzo, st):
        self.pd_gru_gates_bwd(dh, lddh, dh2, lddh2, dh3, lddh3, rzn, ldrzn, hn, ldhn, hp, ldhp, dgi, lddgi, dgh, lddgh,
                              dhp, lddhp, None, 0, lengths, t, B, H, st)
        _arr(zero_out, (B, H), (ldzo, 1))[...] = 0

    def pd_set_pdl(self, on):
        pass

    # ---- pianotree misc ----
    def pd_grid_prepare(self, x, steps, tok, lengths, pt, dt, st):
        X = _arr(x, (steps, 16, 6), (96, 6, 1), np.int64)
        _arr(tok, (steps, 16, 6), (96, 6, 1), np.int32)[...] = X
        if lengths is not None:
            _arr(lengths, (steps,), (1,), np.int32)[...] = 16 - (X[:, :, 0] == 130).sum(-1)
        if pt is not None:
            _arr(pt, (steps, 15), (15, 1), np.int32)[...] = X[:, 1:, 0]
        if dt is not None:
            _arr(dt, (steps, 15, 5), (75, 5, 1), np.int32)[...] = X[:, 1:, 1:]

    def pd_prmat_to_grid(self, pr, n_steps, x, overflow, st):
        P = _arr(pr, (n_steps, 128), (128, 1))
        X = _arr(x, (n_steps, 16, 6), (96, 6, 1), np.int64)
        X[...] = 2
        X[:, :, 0] = 130
        X[:, 0, 0] = 128
        for s in range(n_steps):
            ps = np.nonzero(P[s])[0]
            if len(ps) > 14:
                _arr(overflow, (1,), (1,), np.int32)[0] = 1
                ps = ps[:14]
            for i, p in enumerate(ps):
                d = int(P[s, p]) - 1
                X[s, 1 + i] = [p] + [(d >> (4 - b)) & 1 for b in range(5)]
            X[s, len(ps) + 1, 0] = 129

    def pd_greedy_decode_small(self, B, h_time0, gi_z, wt_tok, ld_wt, wt_hh, bt_hh, init_tok, w_t2n, b_t2n, wn_sum, ld_wn,
                               bn_ih, wn_tok, wn_hh, bn_hh, w_heads, b_heads, d_wih, d_bih, d_whh, d_bhh, d_sos, d_wout,
                               d_bout, emb_wt, emb_b, we_ih_f, we_hh_f, be_ih_f, be_hh_f, we_ih_b, we_hh_b, be_ih_b, be_hh_b,
                               tokens, lens_out, ws, bar, st):
        """numpy restatement of the persistent small-batch greedy decode (ptvae.py:430-491, inference=True)."""
        f = np.float32
        A = lambda ptr, shape, strides=None: _arr(ptr, shape, strides or tuple(
            int(np.prod(shape[i + 1:])) for i in range(len(shape))))
        h_time = A(h_time0, (B, 1024)).copy()
        giz = A(gi_z, (B, 3072))
        Wtt, Wth, bth = A(wt_tok, (3072, 256), (ld_wt, 1)), A(wt_hh, (3072, 1024)), A(bt_hh, (3072,))
        Wt2n, bt2n = A(w_t2n, (512, 1024)), A(b_t2n, (512,))
        Wns, bni = A(wn_sum, (1536, 1024), (ld_wn, 1)), A(bn_ih, (1536,))
        Wnt, Wnh, bnh = A(wn_tok, (1536, 128), (ld_wn, 1)), A(wn_hh, (1536, 512)), A(bn_hh, (1536,))
        Wh, bh = A(w_heads, (194, 512)), A(b_heads, (194,))
        Dwi, Dbi, Dwh, Dbh = A(d_wih, (192, 5)), A(d_bih, (192,)), A(d_whh, (192, 64)), A(d_bhh, (192,))
        Dsos, Dwo, Dbo = A(d_sos, (5,)), A(d_wout, (2, 64)), A(d_bout, (2,))
        WT, eb = A(emb_wt, (135, 128)), A(emb_b, (128,))
        eg = [(A(we_ih_f, (384, 128)), A(we_hh_f, (384, 128)), A(be_ih_f, (384,)), A(be_hh_f, (384,))),
              (A(we_ih_b, (384, 128)), A(we_hh_b, (384, 128)), A(be_ih_b, (384,)), A(be_hh_b, (384,)))]
        TOK = _arr(tokens, (32, 15, B, 6), (15 * B * 6, B * 6, 6, 1), np.int32)
        LO = _arr(lens_out, (32, B), (B, 1), np.int32) if lens_out is not None else None

        def cell(gi, gh, h, H):
            r = _sig(gi[:, :H] + gh[:, :H])
            z = _sig(gi[:, H:2 * H] + gh[:, H:2 * H])
            n = np.tanh(gi[:, 2 * H:] + r * gh[:, 2 * H:], dtype=f)
            return ((1 - z) * n + z * h).astype(f)

        def embed(tok):                                             # tok (B,6) int
            e = np.tile(eb, (B, 1)).astype(f)
            for b in range(B):
                if 0 <= tok[b, 0] < 130:
                    e[b] += WT[tok[b, 0]]
                for k in range(5):
                    e[b] += f(tok[b, 1 + k]) * WT[130 + k]
            return e
        tok_time = np.tile(A(init_tok, (256,)), (B, 1)).astype(f)
        sos = np.tile(np.array([128, 2, 2, 2, 2, 2]), (B, 1))
        for t in range(32):
            h_time = cell(tok_time @ Wtt.T + giz, h_time @ Wth.T + bth, h_time, 1024)
            h_n = (h_time @ Wt2n.T + bt2n).astype(f)
            gi_s = (h_time @ Wns.T + bni).astype(f)
            pred = np.zeros((B, 16, 128), f)
            pred[:, 0] = embed(sos)
            lens = np.zeros(B, np.int64)
            for n in range(1, 16):
                h_n = cell(pred[:, n - 1] @ Wnt.T + gi_s, h_n @ Wnh.T + bnh, h_n, 512)
                heads = (h_n @ Wh.T + bh).astype(f)
                tk = np.zeros((B, 6), np.int64)
                tk[:, 0] = heads[:, :130].argmax(-1)
                hd = heads[:, 130:].copy()
                gi_d = np.tile(Dwi @ Dsos + Dbi, (B, 1)).astype(f)
                for k in range(5):
                    hd = cell(gi_d, hd @ Dwh.T + Dbh, hd, 64)
                    lg = hd @ Dwo.T + Dbo
                    bit = (lg[:, 1] > lg[:, 0]).astype(np.int64)
                    tk[:, 1 + k] = bit
                    gi_d = (Dwi[:, bit].T + Dbi).astype(f)           # one-hot at index == bit value (ptvae.py:322-326)
                TOK[t, n - 1] = tk
                lens = np.where((lens == 0) & (tk[:, 0] == 129), n, lens)
                pred[:, n] = embed(tk)
            lens = np.where(lens == 0, 15, lens)
            if LO is not None:
                LO[t] = lens
            if t == 31:
                break
            out = np.zeros((B, 256), f)
            for d, (wi, wh, bi, bhh) in enumerate(eg):
                for b in range(B):
                    h = np.zeros((1, 128), f)
                    order = range(lens[b] - 1, -1, -1) if d else range(lens[b])
                    for k in order:
                        h = cell((pred[b, k] @ wi.T + bi)[None], h @ wh.T + bhh, h, 128)
                    out[b, d * 128:(d + 1) * 128] = h[0]
            tok_time = out

    def pd_pack_tokens(self, tok, R, out, st):
        T = _arr(tok, (R, 6), (6, 1), np.int32)
        O = _arr(out, (R, 2), (2, 1), np.uint8)
        O[:, 0] = T[:, 0].astype(np.uint8)
        O[:, 1] = sum((T[:, 1 + b] & 1) << (4 - b) for b in range(5)).astype(np.uint8)

    def pd_roll_prmat(self, pr_in, shift, B, pr_out, st):
        P = _arr(pr_in, (B, 32, 128), (4096, 128, 1))
        O = _arr(pr_out, (B, 32, 128), (4096, 128, 1))
        S = _arr(shift, (B,), (1,), np.int32)
        for b in range(B):
            O[b] = np.roll(P[b], int(S[b]), axis=-1)          # converter.py:65-68

    def pd_expand_chord(self, chord14, shift, rows, rows_per_seg, c36, st):
        C = _arr(chord14, (rows, 14), (14, 1))
        O = _arr(c36, (rows, 36), (36, 1))
        S = _arr(shift, ((rows + rows_per_seg - 1) // rows_per_seg,), (1,), np.int32)
        for r in range(rows):                                 # converter.py:150-164
            sh = int(S[r // rows_per_seg])
            root, bass = (int(C[r, 0]) + sh) % 12, (int(C[r, 13]) + sh) % 12
            o = np.zeros(36, np.float32)
            o[root] = 1
            o[12:24] = np.roll(C[r, 1:13], sh)
            o[24 + bass] = 1
            O[r] = o

    def pd_slerp_path(self, z1, z2, B, D, count, out, st):
        A, C = _arr(z1, (B, D), (D, 1)), _arr(z2, (B, D), (D, 1))
        O = _arr(out, (B, count, D), (count * D, D, 1))
        for b in range(B):                                    # model.py:218-242, float64 like the reference
            a, c = A[b].astype(np.float64), C[b].astype(np.float64)
            n1, n2 = np.linalg.norm(a), np.linalg.norm(c)
            u1, u2 = a / n1, c / n2
            ts = np.linspace(0.0, 1.0, count)
            om = np.arccos(np.clip(np.dot(u1, u2), -1.0, 1.0))
            so = np.sin(om)
            dirs = np.sin((1.0 - ts) * om)[:, None] / so * u1[None] + np.sin(ts * om)[:, None] / so * u2[None]
            O[b] = (dirs * np.exp(np.linspace(np.log(n1), np.log(n2), count))[:, None]).astype(np.float32)

    def pd_grid_to_prmat(self, tok, n_steps, pr, st):
        T = _arr(tok, (n_steps, 15, 6), (90, 6, 1), np.int32)
        P = _arr(pr, (n_steps, 128), (128, 1))
        P[...] = 0
        for s in range(n_steps):
            for n in range(10):
                p = T[s, n, 0]
                if p == 129:
                    break
                dur = int("".join(str(int(b)) for b in T[s, n, 1:]), 2) + 1
                if 0 <= p < 128:
                    P[s, p] = min(dur, 32 - s % 32)

    def pd_note_embed_fwd(self, tok, R, WT, bias, out, ldo, st):
        T = _arr(tok, (R, 6), (6, 1), np.int32)
        W = _arr(WT, (135, 128), (128, 1))
        o = np.tile(_arr(bias, (128,), (1,)), (R, 1))
        p = T[:, 0]
        ok = (p >= 0) & (p < 130)
        o[ok] += W[p[ok]]
        o += T[:, 1:].astype(np.float32) @ W[130:]
        _arr(out, (R, 128), (ldo, 1))[...] = o

    def pd_note_embed_bwd(self, tok, R, g, ldg, dWT, db, st):
        T = _arr(tok, (R, 6), (6, 1), np.int32)
        G = _arr(g, (R, 128), (ldg, 1))
        W = _arr(dWT, (135, 128), (128, 1))
        p = T[:, 0]
        ok = (p >= 0) & (p < 130)
        np.add.at(W, p[ok], G[ok])
        W[130:] += T[:, 1:].astype(np.float32).T @ G
        _arr(db, (128,), (1,))[...] += G.sum(0)

    def pd_greedy_pick(self, pitch, ldp, dur, ldd, R, n, tok, ldtok, lens, st):
        P = _arr(pitch, (R, 130), (ldp, 1))
        D = _arr(dur, (R, 5, 2), (ldd, 2, 1))
        T = _arr(tok, (R, 6), (ldtok, 1), np.int32)
        pi = P.argmax(1)
        T[:, 0] = pi
        T[:, 1:] = (D[:, :, 1] > D[:, :, 0]).astype(np.int32)
        if lens is not None:
            L = _arr(lens, (R,), (1,), np.int32)
            L[(L == 0) & (pi == 129)] = n
            if n == 15:
                L[L == 0] = 15

    def pd_greedy_pick_embed(self, pitch, ldp, dur, ldd, R, n, tok, ldtok, lens, WT, bias, emb, lde, st):
        self.pd_greedy_pick(pitch, ldp, dur, ldd, R, n, tok, ldtok, lens, st)
        T = np.ascontiguousarray(_arr(tok, (R, 6), (ldtok, 1), np.int32))
        self.pd_note_embed_fwd(T.ctypes.data, R, WT, bias, emb, lde, st)

    def pd_dur_token(self, logit, ldl, R, tok, st):
        Lg = _arr(logit, (R, 2), (ldl, 1))
        T = _arr(tok, (R, 5), (5, 1))
        one = Lg[:, 1] > Lg[:, 0]
        T[...] = 0
        T[:, 0] = ~one
        T[:, 1] = one

    # ---- fused duration decoder ----
    @staticmethod
    def _dur_tables(w_ih, b_ih, sos):
        W = _arr(w_ih, (192, 5), (5, 1))
        b = _arr(b_ih, (192,), (1,))
        return np.stack([W @ _arr(sos, (5,), (1,)) + b, W[:, 0] + b, W[:, 1] + b])

    def pd_dur_decode_fwd(self, h0, ldh0, Q, w_ih, b_ih, w_hh, b_hh, sos, w_out, b_out, logits, S, tf32, st):
        h = _arr(h0, (Q, 64), (ldh0, 1)).copy()
        gi_t = self._dur_tables(w_ih, b_ih, sos)
        Whh, bhh = _arr(w_hh, (192, 64), (64, 1)), _arr(b_hh, (192,), (1,))
        Wo, bo = _arr(w_out, (2, 64), (64, 1)), _arr(b_out, (2,), (1,))
        L = _arr(logits, (Q, 5, 2), (10, 2, 1))
        Sb = _arr(S, (Q, 6, 72), (432, 72, 1)) if S is not None else None
        tok = np.zeros(Q, np.int64)
        for k in range(5):
            if Sb is not None:
                Sb[:, k, :] = 0
                Sb[:, k, :64] = h
                if k == 0:
                    Sb[:, 0, 64:69] = _arr(sos, (5,), (1,))
                    Sb[:, 0, 70] = 1
                else:
                    Sb[np.arange(Q), k, 64 + tok - 1] = 1
                Sb[:, k, 69] = 1
            gi = gi_t[tok]
            gh = h @ Whh.T + bhh
            r = _sig(gi[:, :64] + gh[:, :64])
            z = _sig(gi[:, 64:128] + gh[:, 64:128])
            n = np.tanh(gi[:, 128:] + r * gh[:, 128:])
            h = ((1 - z) * n + z * h).astype(np.float32)
            lg = h @ Wo.T + bo
            L[:, k, :] = lg
            tok = np.where(lg[:, 1] > lg[:, 0], 2, 1)
        if Sb is not None:
            Sb[:, 5, :] = 0
            Sb[:, 5, :64] = h
            Sb[:, 5, 69] = 1

    def pd_dur_decode_bwd(self, S, dlog, Q, w_ih, b_ih, w_hh, b_hh, sos, w_out, b_out, GX, dh0, lddh0, tf32, st):
        Sb = _arr(S, (Q, 6, 72), (432, 72, 1))
        dL = _arr(dlog, (Q, 5, 2), (10, 2, 1))
        gi_t = self._dur_tables(w_ih, b_ih, sos)
        Whh, bhh = _arr(w_hh, (192, 64), (64, 1)), _arr(b_hh, (192,), (1,))
        Wo = _arr(w_out, (2, 64), (64, 1))
        G = _arr(GX, (Q, 6, 264), (6 * 264, 264, 1))
        G[...] = 0
        dh = np.zeros((Q, 64), np.float32)
        for k in range(4, -1, -1):
            hp = Sb[:, k, :64]
            tok = np.zeros(Q, np.int64) if k == 0 else np.where(Sb[:, k, 65] > 0.5, 2, 1)
            G[:, k + 1, 256:258] = dL[:, k]
            dh = dh + dL[:, k] @ Wo
            gi = gi_t[tok]
            gh = hp @ Whh.T + bhh
            r = _sig(gi[:, :64] + gh[:, :64])
            z = _sig(gi[:, 64:128] + gh[:, 64:128])
            n = np.tanh(gi[:, 128:] + r * gh[:, 128:])
            dn = dh * (1 - z) * (1 - n * n)
            dz = dh * (hp - n) * z * (1 - z)
            dr = dn * gh[:, 128:] * r * (1 - r)
            dgh = np.concatenate([dr, dz, dn * r], 1)
            G[:, k, :192] = dgh
            G[:, k, 192:256] = dn
            dh = (dh * z + dgh @ Whh).astype(np.float32)
        _arr(dh0, (Q, 64), (lddh0, 1))[...] = dh

    def pd_chord_feedback(self, root, ldr, chroma, ldc, bass, ldb, B, flags, tok, ldt, st):
        Rt = _arr(root, (B, 12), (ldr, 1))
        Ch = _arr(chroma, (B, 12, 2), (ldc, 2, 1))
        Bs = _arr(bass, (B, 12), (ldb, 1))
        T = _arr(tok, (B, 36), (ldt, 1))
        fr = np.zeros(12, np.float32)
        fb = np.zeros(12, np.float32)
        fr[Rt.argmax(1)] = 1
        fb[Bs.argmax(1)] = 1
        T[:, :12] = fr
        T[:, 12:24] = (Ch[:, :, 1] > Ch[:, :, 0])
        T[:, 24:] = fb

    def pd_chord_targets(self, c, rows, root, chroma, bass, st):
        C = _arr(c, (rows, 36), (36, 1))
        _arr(root, (rows,), (1,), np.int32)[...] = C[:, :12].argmax(1)
        _arr(chroma, (rows, 12), (12, 1), np.int32)[...] = C[:, 12:24].astype(np.int32)
        _arr(bass, (rows,), (1,), np.int32)[...] = C[:, 24:].argmax(1)

    # ---- texture ----
    @staticmethod
    def _conv(pr, w, b):
        B = pr.shape[0]
        C = w.shape[0]
        bands = pr.reshape(B, 8, 4, 128)
        win = np.lib.stride_tricks.sliding_window_view(bands, 12, axis=3)        # (B,8,4,117,12)
        return np.einsum("bidwk,cdk->bciw", win, w.reshape(C, 4, 12)).astype(np.float32) + b[None, :, None, None]

    def pd_texture_frontend_fwd(self, pr, w, b, B, C, out, st):
        P = _arr(pr, (B, 32, 128), (4096, 128, 1))
        y = np.maximum(self._conv(P, _arr(w, (C, 48), (48, 1)), _arr(b, (C,), (1,))), 0)[..., :116]
        _arr(out, (B, C, 8, 29), (C * 232, 232, 29, 1))[...] = y.reshape(B, C, 8, 29, 4).max(-1)

    def pd_texture_frontend_fwd_ix(self, pr, w, b, B, C, out, amax, st):
        self.pd_texture_frontend_fwd(pr, w, b, B, C, out, st)
        P = _arr(pr, (B, 32, 128), (4096, 128, 1))
        y = self._conv(P, _arr(w, (C, 48), (48, 1)), _arr(b, (C,), (1,)))[..., :116].reshape(B, C, 8, 29, 4)
        q = y.argmax(-1)                                       # first maximum
        active = np.take_along_axis(y, q[..., None], -1)[..., 0] > 0
        _arr(amax, (B, C, 8, 29), (C * 232, 232, 29, 1), np.int8)[...] = np.where(active, q, -1)

    def pd_texture_frontend_bwd_ix(self, pr, amax, B, C, g, dw, db, st):
        P = _arr(pr, (B, 32, 128), (4096, 128, 1))
        A = _arr(amax, (B, C, 8, 29), (C * 232, 232, 29, 1), np.int8).astype(np.int64)
        G = _arr(g, (B, C, 8, 29), (C * 232, 232, 29, 1))
        gy = np.zeros((B, C, 8, 29, 4), np.float32)
        np.put_along_axis(gy, np.maximum(A, 0)[..., None], np.where(A >= 0, G, 0)[..., None], -1)
        gy = gy.reshape(B, C, 8, 116)
        bands = P.reshape(B, 8, 4, 128)
        win = np.lib.stride_tricks.sliding_window_view(bands, 12, axis=3)[:, :, :, :116]   # (B,8,4,116,12)
        _arr(dw, (C, 4, 12), (48, 12, 1))[...] += np.einsum("bciw,bidwk->cdk", gy, win)
        _arr(db, (C,), (1,))[...] += gy.sum((0, 2, 3))

    def pd_texture_frontend_bwd(self, pr, w, b, B, C, g, dw, db, st):
        P = _arr(pr, (B, 32, 128), (4096, 128, 1))
        y = self._conv(P, _arr(w, (C, 48), (48, 1)), _arr(b, (C,), (1,)))[..., :116].reshape(B, C, 8, 29, 4)
        G = _arr(g, (B, C, 8, 29), (C * 232, 232, 29, 1))
        q = y.argmax(-1)
        active = np.take_along_axis(y, q[..., None], -1)[..., 0] > 0
        gy = np.zeros_like(y)
        np.put_along_axis(gy, q[..., None], np.where(active, G, 0)[..., None], -1)
        gy = gy.reshape(B, C, 8, 116)
        bands = P.reshape(B, 8, 4, 128)
        win = np.lib.stride_tricks.sliding_window_view(bands, 12, axis=3)[:, :, :, :116]   # (B,8,4,116,12)
        _arr(dw, (C, 4, 12), (48, 12, 1))[...] += np.einsum("bciw,bidwk->cdk", gy, win)
        _arr(db, (C,), (1,))[...] += gy.sum((0, 2, 3))

    # ---- losses ----
    @staticmethod
    def _lse(L):
        m = L.max(1, keepdims=True)
        return m[:, 0] + np.log(np.exp(L - m).sum(1))

    def pd_ce_fwd(self, logits, ldl, tgt, R, C, ignore, acc, loss, st):
        L = _arr(logits, (R, C), (ldl, 1))
        T = _arr(tgt, (R,), (1,), np.int32)
        ok = T != ignore
        A = _arr(acc, (2,), (1,))
        A[0] = (self._lse(L[ok]) - L[ok, T[ok]]).sum() if ok.any() else 0.0
        A[1] = ok.sum()
        _arr(loss, (1,), (1,))[0] = A[0] / A[1]

    def pd_ce_bwd(self, logits, ldl, tgt, R, C, ignore, acc, gout, dl, lddl, st):
        L = _arr(logits, (R, C), (ldl, 1))
        T = _arr(tgt, (R,), (1,), np.int32)
        ok = T != ignore
        scale = _arr(gout, (1,), (1,))[0] / _arr(acc, (2,), (1,))[1]
        m = L.max(1, keepdims=True)
        p = np.exp(L - m)
        p /= p.sum(1, keepdims=True)
        p[np.arange(R)[ok], T[ok]] -= 1
        _arr(dl, (R, C), (lddl, 1))[...] = np.where(ok[:, None], p * scale, 0)

    def pd_exp_fwd(self, x, n, y, st):
        _arr(y, (n,), (1,))[...] = np.exp(_arr(x, (n,), (1,)))

    def pd_mul_f32(self, a, b, n, out, st):
        _arr(out, (n,), (1,))[...] = _arr(a, (n,), (1,)) * _arr(b, (n,), (1,))

    def pd_select_rows(self, a, lda, b, ldb, flag, out, ldo, rows, cols, st):
        f = int(_arr(flag, (1,), (1,), np.int32)[0])
        _arr(out, (rows, cols), (ldo, 1))[...] = _arr(a, (rows, cols), (lda, 1)) if f else _arr(b, (rows, cols), (ldb, 1))

    def pd_select_rows_bwd(self, dout, ldd, flag, da, ldda, db, lddb, rows, cols, st):
        f = int(_arr(flag, (1,), (1,), np.int32)[0])
        g = _arr(dout, (rows, cols), (ldd, 1))
        if da is not None:
            _arr(da, (rows, cols), (ldda, 1))[...] = g if f else 0.0
        if db is not None:
            _arr(db, (rows, cols), (lddb, 1))[...] = 0.0 if f else g

    def pd_add_f32(self, a, b, n, out, st):
        _arr(out, (n,), (1,))[...] = _arr(a, (n,), (1,)) + _arr(b, (n,), (1,))

    def pd_reparam_fwd(self, mu, sd, eps, B, D, z, ldz, st):
        m = _arr(mu, (B, D), (D, 1))
        _arr(z, (B, D), (ldz, 1))[...] = m if eps is None else m + _arr(sd, (B, D), (D, 1)) * _arr(eps, (B, D), (D, 1))

    def pd_reparam_bwd(self, dz, lddz, eps, B, D, dmu, dsd, st):
        g = _arr(dz, (B, D), (lddz, 1))
        _arr(dmu, (B, D), (D, 1))[...] = g
        _arr(dsd, (B, D), (D, 1))[...] = 0 if eps is None else g * _arr(eps, (B, D), (D, 1))

    def pd_kl_fwd(self, mu, sd, n, out, st):
        m, s = _arr(mu, (n,), (1,)), _arr(sd, (n,), (1,))
        _arr(out, (1,), (1,))[0] = (-np.log(s) + 0.5 * (s * s + m * m) - 0.5).mean()

    def pd_kl_bwd(self, mu, sd, n, gout, dmu, dsd, st):
        m, s = _arr(mu, (n,), (1,)), _arr(sd, (n,), (1,))
        g = _arr(gout, (1,), (1,))[0] / n
        _arr(dmu, (n,), (1,))[...] = g * m
        _arr(dsd, (n,), (1,))[...] = g * (s - 1 / s)


    # ---- optimizer tail ----
    def pd_sumsq_f32(self, g, n, out, st):
        G = _arr(g, (n,), (1,))
        _arr(out, (1,), (1,))[0] += np.float32((G.astype(np.float64) ** 2).sum())

    def pd_counter_inc(self, c, st):
        _arr(c, (1,), (1,), np.int32)[0] += 1

    def pd_adam_clip_step(self, p, g, m, v, n, sumsq, step, lr0, gamma, lr_min, b1, b2, eps, clip, st):
        P, G, M, V = (_arr(x, (n,), (1,)) for x in (p, g, m, v))
        t = int(_arr(step, (1,), (1,), np.int32)[0])
        coef = min(1.0, clip / (np.sqrt(_arr(sumsq, (1,), (1,))[0]) + 1e-6)) if clip > 0 else 1.0
        lr = max(lr0 * gamma ** (t - 1), lr_min) if gamma > 0 else lr0
        gi = G * np.float32(coef)
        M[...] = b1 * M + (1 - b1) * gi
        V[...] = b2 * V + (1 - b2) * gi * gi
        P[...] -= (lr / (1 - b1 ** t)) * M / (np.sqrt(V) / np.sqrt(1 - b2 ** t) + eps)


def poison_empty(monkeypatch):
    """Fill every floating-point ``torch.empty`` with NaN for the duration of a test: reads of rows that a kernel
    legitimately leaves unwritten (dead rows of the packed note level) then poison whatever they leak into."""
    import torch
    real = torch.empty

    def empty(*a, **k):
        t = real(*a, **k)
        if t.is_floating_point():
            t.fill_(float("nan"))
        return t
    monkeypatch.setattr(torch, "empty", empty)


def install(monkeypatch):
    """Route polydis_b200's library calls to the numpy emulation for the duration of a test."""
    from polydis_b200 import _lib, ops
    be = CpuBackend()
    monkeypatch.setattr(_lib, "call", be.call)
    monkeypatch.setattr(ops, "_call", be.call)
    monkeypatch.setattr(ops, "_stream", lambda: None)
    monkeypatch.setattr(ops, "_chk", lambda t, name="tensor": t)
    return be
