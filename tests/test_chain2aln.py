"""Round-flattened chain -> alignment driver: every seed extended speculatively in
one GPU launch sequence + host replay of testExtension / checkOverlapping / MARKED / seed coverage must give
the region lists of the reference's round loop (oracle: extensions on demand, in the reference's order)."""
import numpy as np
import pytest


def _workload_chains(pkg, seed=11, n_pairs=600, L=151, G=400000, eps=0.03):
    W = pkg.workload
    opt = pkg.jni.MemOptType()
    rng = np.random.default_rng(seed)
    ref = W.make_reference(G, seed)
    rb = W.ReadBatch(ref, n_pairs, L, eps, 400, 50, rng, indel_frac=0.5)
    rco, chains, seeds = W.all_seeds(rb, opt)
    return opt, ref, rb.reads, rco, chains, seeds


def _crafted_chains(pkg, rng, ref, L=101, n_reads=400):
    """Adversarial chain sets: several chains per read, seeds overlapping on different diagonals, duplicates,
    seeds spanning the read, reads without chains, reverse-strand chains (coordinates >= l_pac)."""
    G = len(ref)
    comp = pkg.workload.COMP
    reads = np.zeros((n_reads, L), dtype=np.uint8)
    rco, chains, seeds = [0], [], []
    for r in range(n_reads):
        pos = int(rng.integers(2000, G - 2000))
        rev = rng.random() < 0.3
        frag = ref[pos:pos + L].copy()
        k = int(rng.integers(0, 4))
        for _ in range(k):                                        # substitutions
            frag[int(rng.integers(0, L))] = rng.integers(0, 4)
        read = comp[frag[::-1]] if rev else frag
        reads[r] = read
        base = (2 * G - (pos + L)) if rev else pos               # doubled coordinate of read base 0
        n_ch = int(rng.choice([0, 1, 1, 1, 2, 3]))
        for _c in range(n_ch):
            ns = int(rng.choice([1, 1, 2, 3, 5, 8]))
            first = len(seeds)
            qs = sorted(int(x) for x in rng.integers(0, L - 19, ns))
            for q in qs:
                ln = int(min(L - q, rng.choice([19, 20, 25, 40, 60, L])))
                dg = int(rng.choice([0, 0, 0, 1, -1, 3, -7, 20]))        # diagonal shift: overlapping seeds off the main diagonal
                seeds.append((base + q + dg, q, ln))
            chains.append((first, ns))
        rco.append(len(chains))
    return (reads, np.array(rco, dtype=np.int32), np.array(chains, dtype=pkg._lib.CHAIN_DTYPE).reshape(-1),
            np.array(seeds, dtype=pkg._lib.SEED_DTYPE).reshape(-1))


def test_oracle_round_loop_sanity(pkg, oracle):
    """CPU only: the restated round loop on workload chains -- one region per read for clean reads, the extra
    seeds of a read skipped (MARKED) once the first extension covers them."""
    opt, ref, reads, rco, chains, seeds = _workload_chains(pkg, n_pairs=150, eps=0.02)
    pac = pkg.jni.packPac(ref)
    regs, off, cells, n_ext = oracle.chain2aln(reads, rco, chains, seeds, pac, len(ref))
    assert off[-1] == len(regs) and len(regs) >= int((np.diff(rco) > 0).sum())
    assert n_ext <= len(seeds) and n_ext >= len(regs) * 0.5 and cells > 0
    per_read = np.diff(off)
    assert (per_read[np.diff(rco) > 0] >= 1).all() and per_read.mean() < 1.6
    assert (regs["seedcov"] >= 19).all() and (regs["qe"] > regs["qb"]).all()


def test_oracle_pinned_vs_reference_c(pkg, oracle):
    """orc_chain2aln (restatement of the Scala round loop) against the reference's own bwa-0.7.8
    mem_chain2aln compiled from /root/reference (oracle/_ref/libbwamem_ref.so): identical region
    lists with zdrop = 0, where the Scala and the C z-drop rules coincide, on workload
    chains and on crafted adversarial chain sets, both strands; with the default zdrop the two may
    differ only through that documented quirk."""
    if not oracle.ref_mem_available():
        pytest.skip("oracle/_ref/libbwamem_ref.so not built")
    opt, ref, reads, rco, chains, seeds = _workload_chains(pkg, n_pairs=200, eps=0.03)
    pac = pkg.jni.packPac(ref)
    o0 = oracle.default_opt(); o0.zdrop = 0
    for (rd, rc_, ch, sd) in ((reads, rco, chains, seeds), _crafted_chains(pkg, np.random.default_rng(14), ref, n_reads=300)):
        want, woff = oracle.ref_mem_chain2aln(rd, rc_, ch, sd, pac, len(ref), zdrop=0)
        got, goff, _, _ = oracle.chain2aln(rd, rc_, ch, sd, pac, len(ref), opt=o0)
        assert np.array_equal(goff, woff)
        assert got.tobytes() == want.tobytes()
        want, woff = oracle.ref_mem_chain2aln(rd, rc_, ch, sd, pac, len(ref))                # default zdrop = 100
        got, goff, _, _ = oracle.chain2aln(rd, rc_, ch, sd, pac, len(ref))
        same = np.array_equal(goff, woff) and got.tobytes() == want.tobytes()
        assert same or (np.array_equal(goff, woff) and (got != want).mean() < 0.02)


def test_oracle_pinned_vs_reference_c_default_zdrop(pkg, oracle):
    """The same pin at the DEFAULT zdrop = 100, read by read: the oracle counts the SWExtend rows in which the Scala
    z-drop rule and the C's decide differently from the same state (orc_zdrop_divergences); for every read whose
    extensions met no such row the two round loops are the same computation, and the region list must be
    byte-identical to a run of the reference's mem_chain2aln with its default options.  Indel-rich and high-error
    reads so that the z-drop tests run (and fire) in many rows."""
    if not oracle.ref_mem_available():
        pytest.skip("oracle/_ref/libbwamem_ref.so not built")
    n_same = n_div = 0
    for eps, seed in ((0.03, 21), (0.08, 22)):
        opt, ref, reads, rco, chains, seeds = _workload_chains(pkg, seed=seed, n_pairs=150, eps=eps)
        pac = pkg.jni.packPac(ref)
        sets = [(reads, rco, chains, seeds), _crafted_chains(pkg, np.random.default_rng(seed), ref, n_reads=150)]
        for rd, rc_, ch, sd in sets:
            for r in range(len(rd)):
                c0, c1 = int(rc_[r]), int(rc_[r + 1])
                if c0 == c1:
                    continue
                one_rco = np.array([0, c1 - c0], dtype=np.int32)
                oracle.zdrop_divergences(reset=True)
                got, goff, _, _ = oracle.chain2aln(rd[r:r + 1], one_rco, ch[c0:c1], sd, pac, len(ref))
                if oracle.zdrop_divergences() != 0:
                    n_div += 1
                    continue
                want, woff = oracle.ref_mem_chain2aln(rd[r:r + 1], one_rco, ch[c0:c1], sd, pac, len(ref))
                assert np.array_equal(goff, woff), r
                assert got.tobytes() == want.tobytes(), r
                n_same += 1
    assert n_same > 500 and n_div < n_same


def test_oracle_round_loop_equals_reference_c_with_its_zdrop_rule(pkg, oracle):
    """The restated round loop with ONLY the z-drop decision swapped for the C's (oracle.c_zdrop_rule): byte-identical
    to the reference's mem_chain2aln with its default options on every read -- so that decision is all that separates
    orc_chain2aln from a run of the reference's C."""
    if not oracle.ref_mem_available():
        pytest.skip("oracle/_ref/libbwamem_ref.so not built")
    opt, ref, reads, rco, chains, seeds = _workload_chains(pkg, seed=23, n_pairs=200, eps=0.08)
    pac = pkg.jni.packPac(ref)
    for rd, rc_, ch, sd in ((reads, rco, chains, seeds), _crafted_chains(pkg, np.random.default_rng(23), ref, n_reads=300)):
        want, woff = oracle.ref_mem_chain2aln(rd, rc_, ch, sd, pac, len(ref))
        with oracle.c_zdrop_rule():
            got, goff, _, _ = oracle.chain2aln(rd, rc_, ch, sd, pac, len(ref))
        assert np.array_equal(goff, woff)
        assert got.tobytes() == want.tobytes()


@pytest.mark.gpu
def test_chain2aln_flat_parity(pkg, oracle):
    L_ = pkg.lib()
    assert L_.csbwa_init(0) >= 1
    # (1) workload chains (indel-rich so that seeds sit on different diagonals)
    opt, ref, reads, rco, chains, seeds = _workload_chains(pkg)
    pac = pkg.jni.packPac(ref)
    pkg.jni.refUpload(pac, len(ref))
    want, woff, cells, n_ext = oracle.chain2aln(reads, rco, chains, seeds, pac, len(ref))
    before = pkg.stats()["ext_cells"]
    got, goff, n_spec, n_used = pkg.jni.memChainToAlnBatched(reads, rco, chains, seeds, opt, device=0)
    assert np.array_equal(goff, woff)
    assert got.tobytes() == want.tobytes()
    assert n_used == n_ext and n_spec >= n_used                    # consumed == what the reference would run
    assert pkg.stats()["ext_cells"] - before >= cells              # speculation only ever adds work
    # (2) crafted adversarial chains, both strands
    rng = np.random.default_rng(12)
    creads, crco, cchains, cseeds = _crafted_chains(pkg, rng, ref)
    want, woff, _, n_ext = oracle.chain2aln(creads, crco, cchains, cseeds, pac, len(ref))
    got, goff, n_spec, n_used = pkg.jni.memChainToAlnBatched(creads, crco, cchains, cseeds, opt, device=0)
    assert np.array_equal(goff, woff) and got.tobytes() == want.tobytes()
    assert n_used == n_ext and n_spec > n_used                      # some seeds really are skipped
    assert (np.diff(crco) == 0).any() and (np.diff(crco) >= 2).any()
    assert (cseeds["r_beg"] >= len(ref)).any()
    L_.csbwa_ref_release(-1)


@pytest.mark.gpu
def test_human_sized_reference_coordinates(pkg, oracle):
    """BASELINE config 4 shape: a 3.1 Gbp reference (775 MB of .pac, replicated per GPU), seeds whose
    doubled coordinates exceed 2^32 on both strands; coordinate tasks and the flattened driver must
    still agree with the oracle (64-bit addressing of the resident reference)."""
    L_ = pkg.lib()
    assert L_.csbwa_init(0) >= 1
    rng = np.random.default_rng(13)
    l_pac = 3_100_000_000
    pac = rng.integers(0, 256, size=(l_pac + 3) // 4, dtype=np.uint8)
    comp = pkg.workload.COMP
    Lr, n = 151, 256
    reads = np.zeros((n, Lr), dtype=np.uint8)
    seeds = np.zeros(n, dtype=pkg._lib.SEED_DTYPE)
    for r in range(n):
        # positions spread over the whole forward strand, every other read from the reverse strand
        pos = int(rng.integers(1000, l_pac - 1000)) if r % 4 else l_pac - 2000 - r      # also right below the strand boundary
        k = np.arange(pos, pos + Lr, dtype=np.int64)
        frag = ((pac[k >> 2] >> ((~k & 3) << 1)) & 3).astype(np.uint8)
        q = int(rng.integers(0, 60)); ln = int(rng.integers(19, 70))
        for _ in range(2):
            frag_m = frag
            j = int(rng.integers(0, Lr))
            if not (q <= j < q + ln):
                frag_m = frag.copy(); frag_m[j] = (frag[j] + 1) & 3; frag = frag_m
        rev = r % 2 == 1
        reads[r] = comp[frag[::-1]] if rev else frag
        base = (2 * l_pac - (pos + Lr)) if rev else pos
        qq = (Lr - q - ln) if rev else q
        seeds[r] = (base + qq, qq, ln)
    chains = np.zeros(n, dtype=pkg._lib.CHAIN_DTYPE)
    chains["seed_off"] = np.arange(n); chains["n_seeds"] = 1
    rco = np.arange(n + 1, dtype=np.int32)
    assert (seeds["r_beg"] > 2 ** 32).any() and (seeds["r_beg"] < l_pac).any()
    want, woff, _, n_ext = oracle.chain2aln(reads, rco, chains, seeds, pac, l_pac)
    pkg.jni.refUpload(pac, l_pac, device=0)
    got, goff, n_spec, n_used = pkg.jni.memChainToAlnBatched(reads, rco, chains, seeds, device=0)
    assert np.array_equal(goff, woff) and got.tobytes() == want.tobytes()
    assert n_used == n_ext == n and (got["score"] > 100).mean() > 0.9          # the reads really align at those coordinates
    L_.csbwa_ref_release(-1)
