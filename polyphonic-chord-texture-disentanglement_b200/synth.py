"""Seeded synthetic 2-bar segments for benchmarks and parity tests (host side, numpy).

There is no dataset on the GPU box (the reference's ``data/`` is un-shipped, dataset.py:13-14), so
every bench / test input comes from here.  The three tensors are exactly what the reference's
loader hands to the model (dataset_loaders.py:28-34):

* ``pr_mat`` (B,32,128) fp32 -- duration (in 16th steps) at the onset cell, else 0
  (what converter.py:87-113 ``piano_roll_to_target`` yields);
* ``x``      (B,32,16,6) int64 -- the PianoTree grid built from ``pr_mat`` the way
  converter.py:116-147 ``target_to_3dtarget`` is called in dataset.py:98-104
  (max_note_count=16, pitch SOS/EOS/PAD = 128/129/130, dur PAD = 2, 5 duration bits MSB first);
* ``c``      (B,8,36) fp32 -- root one-hot(12) + chroma bits(12) + bass one-hot(12)
  (converter.py:150-164 ``expand_chord`` with shift 0).

``pr_mat_to_grid`` is a vectorised restatement of ``target_to_3dtarget`` (tests/ pin it against the
golden vectors made from the reference's own converter).
"""
import numpy as np

MAX_SIMU_NOTE = 16
PITCH_SOS, PITCH_EOS, PITCH_PAD, DUR_PAD = 128, 129, 130, 2


def pr_mat_to_grid(pr_mat):
    """(B,32,128) durations-at-onset -> (B,32,16,6) int64 PianoTree grid.

    Same result as mapping converter.py:116-147 over the batch with the dataset.py:98-104 arguments.
    Raises if a step holds more than 14 notes (the reference writes out of bounds there).
    """
    pr = np.asarray(pr_mat)
    assert pr.ndim == 3 and pr.shape[1:] == (32, 128)
    B = pr.shape[0]
    on = pr != 0
    cnt = on.sum(-1)                                   # (B,32)
    if cnt.size and cnt.max() > MAX_SIMU_NOTE - 2:
        raise ValueError("more than 14 simultaneous onsets do not fit the 16-slot grid")
    grid = np.full((B, 32, MAX_SIMU_NOTE, 6), DUR_PAD, dtype=np.int64)
    grid[..., 0] = PITCH_PAD
    grid[:, :, 0, 0] = PITCH_SOS
    b, t, p = np.nonzero(on)                           # row-major: pitches ascending inside (b,t)
    slot = np.cumsum(on, axis=-1)[b, t, p]             # 1..k
    grid[b, t, slot, 0] = p
    d = pr[b, t, p].astype(np.int64) - 1
    for k in range(5):
        grid[b, t, slot, 1 + k] = (d >> (4 - k)) & 1
    bb, tt = np.meshgrid(np.arange(B), np.arange(32), indexing="ij")
    grid[bb, tt, cnt + 1, 0] = PITCH_EOS
    return grid


def synth_batch(B, seed=0, max_notes=8, p_active=0.6, lo=36, hi=96):
    """Return ``x (B,32,16,6) int64, c (B,8,36) float32, pr_mat (B,32,128) float32`` (numpy).

    Per (segment, step): active with prob ``p_active``; k ~ U{1..max_notes} distinct pitches from
    [lo,hi); each duration ~ U{1..32-t}.  Per chord slot: root ~ U{0..11}, 12 chroma bits
    ~ Bernoulli(0.3), bass ~ U{0..11}.  Deterministic in (B, seed, ...).

    This is the VECTORISED form of the recipe (one PCG64 stream, whole-batch draws: 65,536 segments in about a second);
    ``synth_batch_recipe`` below is the recipe of BASELINE.md / SURVEY.md appendix A verbatim (a ``RandomState`` loop over
    segments and steps): the same distribution, a different random stream.  The golden fixtures and every test / bench
    input are drawn with this function.
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    n_p = hi - lo
    active = rng.random((B, 32)) < p_active
    k = rng.integers(1, max_notes + 1, size=(B, 32))
    k = np.where(active, k, 0)
    score = rng.random((B, 32, n_p))
    rank = np.argsort(np.argsort(score, axis=-1), axis=-1)      # rank of each pitch
    chosen = rank < k[..., None]
    max_d = 32 - np.arange(32)
    dur = rng.integers(1, max_d[None, :, None] + 1, size=(B, 32, n_p))
    pr_mat = np.zeros((B, 32, 128), dtype=np.float32)
    pr_mat[:, :, lo:hi] = np.where(chosen, dur, 0)
    x = pr_mat_to_grid(pr_mat)

    c = np.zeros((B, 8, 36), dtype=np.float32)
    root = rng.integers(0, 12, size=(B, 8))
    bass = rng.integers(0, 12, size=(B, 8))
    chroma = rng.random((B, 8, 12)) < 0.3
    bi, si = np.meshgrid(np.arange(B), np.arange(8), indexing="ij")
    c[bi, si, root] = 1.0
    c[:, :, 12:24] = chroma
    c[bi, si, 24 + bass] = 1.0
    return x, c, pr_mat


def synth_batch_recipe(B, seed=0, max_notes=8):
    """The synthetic-input recipe of BASELINE.md section 3 / SURVEY.md appendix A, statement for statement: one
    ``np.random.RandomState(seed)``, a python loop over segments, steps and chord slots.  Same distribution as
    ``synth_batch`` (tests/test_api_surface_cpu.py compares their statistics); use it to reproduce the survey's probe
    numbers, not for large batches."""
    rng = np.random.RandomState(seed)
    pr_mat = np.zeros((B, 32, 128), dtype=np.float32)
    c = np.zeros((B, 8, 36), dtype=np.float32)
    for b in range(B):
        for t in range(32):
            if rng.rand() < 0.6:
                k = rng.randint(1, max_notes + 1)
                ps = rng.choice(np.arange(36, 96), size=k, replace=False)
                pr_mat[b, t, ps] = rng.randint(1, min(32, 32 - t) + 1, size=k)
        for s_ in range(8):
            root = rng.randint(12)
            chroma = rng.rand(12) < 0.3
            bass = rng.randint(12)
            c[b, s_, root] = 1.0                 # expand_chord(., shift 0): root one-hot | chroma bits | bass one-hot
            c[b, s_, 12:24] = chroma
            c[b, s_, 24 + bass] = 1.0
    return pr_mat_to_grid(pr_mat), c, pr_mat
