"""Seeded synthetic tasks shared by the parity tests (no reference code involved)."""
import numpy as np

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def rand_seq(rng, n, alphabet=ACGT):
    if n == 0:
        return b""
    return alphabet[rng.integers(0, len(alphabet), n)].tobytes()


def mutate(rng, seq, sub=0.02, indel=0.005, n_rate=0.0):
    out = bytearray()
    for c in seq:
        r = rng.random()
        if r < indel / 2:
            continue  # deletion
        if r < indel:
            out.append(int(ACGT[rng.integers(0, 4)]))  # insertion before
        if rng.random() < sub:
            c = int(ACGT[rng.integers(0, 4)])
        if n_rate and rng.random() < n_rate:
            c = ord("N")
        out.append(c)
    return bytes(out)


def planted_split_cluster(rng, R1, R2):
    """Two windows and a function that draws reads spanning (or not) the planted junction."""
    ref1 = rand_seq(rng, R1)
    ref2 = rand_seq(rng, R2)
    bp1 = int(rng.integers(R1 // 3, R1 - 5)) if R1 > 20 else R1
    bp2 = int(rng.integers(5, 2 * R2 // 3)) if R2 > 20 else 0
    fusion = ref1[:bp1] + ref2[bp2:]

    def draw(L, kind):
        if kind == "span" and len(fusion) > L:
            lo = max(0, bp1 - L + 4)
            hi = min(bp1 - 4, len(fusion) - L)
            if hi < lo:
                lo, hi = 0, max(0, len(fusion) - L)
            s = int(rng.integers(lo, hi + 1))
            return fusion[s:s + L]
        if kind == "left" and R1 >= L:
            s = int(rng.integers(0, R1 - L + 1))
            return ref1[s:s + L]
        if kind == "right" and R2 >= L:
            s = int(rng.integers(0, R2 - L + 1))
            return ref2[s:s + L]
        return rand_seq(rng, L)

    return ref1, ref2, draw


def split_batch(rng, n_clusters, reads_per_cluster, L, R_lo, R_hi, sub=0.01, indel=0.002, n_rate=0.002,
                kinds=("span", "span", "left", "right", "random")):
    refs, reads, task_cluster, task_read = [], [], [], []
    for c in range(n_clusters):
        R1 = int(rng.integers(R_lo, R_hi + 1))
        R2 = int(rng.integers(R_lo, R_hi + 1))
        ref1, ref2, draw = planted_split_cluster(rng, R1, R2)
        refs += [ref1, ref2]
        for _ in range(reads_per_cluster):
            Lr = L if isinstance(L, int) else int(rng.integers(L[0], L[1] + 1))
            read = mutate(rng, draw(Lr, kinds[int(rng.integers(0, len(kinds)))]), sub, indel, n_rate)
            task_cluster.append(c)
            task_read.append(len(reads))
            reads.append(read)
    return refs, reads, np.array(task_cluster, np.int32), np.array(task_read, np.int32)


def simple_batch(rng, n_refs, n_tasks, R, L, related=0.8, sub=0.02, indel=0.002):
    refs = [rand_seq(rng, R if isinstance(R, int) else int(rng.integers(R[0], R[1] + 1))) for _ in range(n_refs)]
    seqs, task_ref, task_seq = [], [], []
    for t in range(n_tasks):
        r = int(rng.integers(0, n_refs))
        Lr = L if isinstance(L, int) else int(rng.integers(L[0], L[1] + 1))
        ref = refs[r]
        if rng.random() < related and len(ref) >= Lr:
            s = int(rng.integers(0, len(ref) - Lr + 1))
            seq = mutate(rng, ref[s:s + Lr], sub, indel)
        else:
            seq = rand_seq(rng, Lr)
        task_ref.append(r)
        task_seq.append(len(seqs))
        seqs.append(seq)
    return refs, seqs, np.array(task_ref, np.int32), np.array(task_seq, np.int32)
