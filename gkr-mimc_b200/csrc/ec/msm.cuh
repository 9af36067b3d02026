// G1Affine.MultiExp and G2Affine.MultiExp on the device: the bucket method (Pippenger) laid out for a GPU, written once over the
// curve (curve.cuh: G1 over Fp, G2 over Fp2).
//
// Replaces gnark-crypto's ecc/bn254 G1Affine.MultiExp / G2Affine.MultiExp (reference go.mod:7, un-vendored; a goroutine per
// window, each walking all scalars) behind their call sites prover/gadget/hints.go:182-183 (InitialRandomnessHint: the 3N GKR
// inputs/outputs against pubKGkr / privKGkrSigma), prover/gadget/prove.go:76,91,189,202,221 (G1) and prove.go:277 (G2, Bs).
// result = sum_i scalars[i] * points[i].
//
// Plan (all sizes from `MsmPlan`): scalars are cut into W signed c-bit digits d in [-2^(c-1), 2^(c-1)]; a non-zero digit sends
// +-point i to bucket |d| - 1 of window w.  Instead of W passes over the input with per-window bucket arrays, every (window,
// bucket) pair is one KEY and the input is counting-sorted by key once:
//   1. k_count    one thread per scalar: (Montgomery -> regular if asked), digits, histogram of keys (L2 atomics)
//   2. scans      exclusive prefix sums of the histogram -> entry offsets; of ceil(count / T) -> task offsets
//   3. k_scatter  one thread per scalar: digits again, entries[off[key] + cursor[key]++] = i | sign << 31
//   4. k_accum    one thread per TASK = at most T consecutive entries of one key: XYZZ sum of the gathered affine points
//                 (madd-2008-s, 8 products + 2 squarings each) -- this is where the work is: n * W mixed additions.  Tasks, not
//                 buckets, are the unit of parallelism, so a skewed input (the reference benchmark hashes ONE value 2^k times:
//                 all scalars equal, one bucket per window holds everything) still spreads over the whole GPU.
//   5. k_group    one thread per GROUP of at most 32 partial sums of one key; k_bucket: one thread per key: sum of its group sums
//                 (one partial sum and one group, usually; thousands of partial sums for the hot buckets of a skewed input)
//   6. k_chunk    per window, per chunk of L consecutive buckets: sum_b (b + 1) * S_b by the running-sum trick on the chunk plus
//                 (first bucket index) * (chunk total)
//   7. k_window   one thread per window: sum of its chunks;  k_final: one thread: Horner over the windows (c doublings each),
//                 affine conversion (one inversion), result in Montgomery and in regular form.
// Every kernel is "one thread = one index", no shared memory and no intra-block cooperation: the bodies below are plain
// functions of the index, launched through an executor (`Exec`).  The CUDA executor (ec.cu) runs them as grids of 128-thread
// blocks; tests/emu/ec_emu.cpp runs the SAME bodies and the same driver loop by loop on the CPU against the oracle.  Entry
// order inside a bucket depends on the atomics and differs from run to run; the group sum, and therefore every output byte,
// does not.
#pragma once
#include "curve.cuh"

namespace ec {

struct MsmPlan {
    uint32_t n;             // points / scalars
    uint32_t c;             // window width in bits, 2..16
    uint32_t W;             // windows: floor(254 / c) + 1, so that W * c >= 255 and the top digit never carries out
    uint32_t B;             // buckets per window = 2^(c-1)
    uint32_t nkeys;         // W * B
    uint32_t T;             // entries per accumulation task
    uint32_t L;             // buckets per reduction chunk
    uint32_t nchunks;       // ceil(B / L)
    uint32_t scalars_mont;  // 1: scalars arrive in Montgomery form (fr.Element memory); 0: regular form (what MultiExp takes)
    uint32_t scan_L;        // elements per scan thread
    uint64_t max_tasks;     // upper bound of the number of accumulation tasks
    uint32_t G;             // partial sums per reduction group (second level of the bucket reduction)
    uint64_t max_groups;    // upper bound of the number of groups
};

// cost model: W * (n mixed additions + ~3 mixed-addition equivalents per bucket for the reduction)
inline MsmPlan msm_make_plan(size_t n, int scalars_mont, int c_force = 0, int T_force = 0) {
    MsmPlan pl;
    pl.n = (uint32_t)n;
    uint32_t best_c = 2;
    double best = 1e300;
    for (uint32_t c = 2; c <= 16; c++) {
        const uint32_t W = 254 / c + 1;
        const double cost = (double)W * ((double)n + 3.0 * (double)(1u << (c - 1)));
        if (cost < best) best = cost, best_c = c;
    }
    pl.c = (c_force >= 2 && c_force <= 16) ? (uint32_t)c_force : best_c;
    pl.W = 254 / pl.c + 1;
    pl.B = 1u << (pl.c - 1);
    pl.nkeys = pl.W * pl.B;
    // task size: twice the mean bucket load, so that a balanced input has one task per bucket with lengths near T / 2 (threads
    // of a warp finish together) and an oversized bucket is cut into pieces no longer than that
    uint64_t mean = ((uint64_t)n + pl.B - 1) / pl.B;
    uint64_t T = 2 * mean;
    if (T < 16) T = 16;
    if (T > 4096) T = 4096;
    pl.T = T_force > 0 ? (uint32_t)T_force : (uint32_t)T;
    uint32_t lg = 0;
    while ((1u << (2 * lg)) < pl.B) lg++;  // L ~ sqrt(B)
    pl.L = 1u << lg;
    pl.nchunks = (pl.B + pl.L - 1) / pl.L;
    pl.scalars_mont = scalars_mont ? 1 : 0;
    pl.scan_L = 256;
    pl.max_tasks = ((uint64_t)n * pl.W) / pl.T + pl.nkeys;
    pl.G = 32;
    pl.max_groups = pl.max_tasks / pl.G + pl.nkeys;
    return pl;
}

// ---- per-thread helpers -------------------------------------------------------------------------------------------------------
EC_HD uint32_t ec_atomic_inc(uint32_t* p) {
#ifdef __CUDA_ARCH__
    return atomicAdd(p, 1u);
#else
    return (*p)++;
#endif
}
EC_HD void ec_flag(uint32_t* p, uint32_t bit) {
#ifdef __CUDA_ARCH__
    atomicOr(p, bit);
#else
    *p |= bit;
#endif
}
enum : uint32_t { MSM_ERR_SCALAR_RANGE = 1u, MSM_ERR_INTERNAL = 2u };

// scalar i in regular form; a non-canonical regular-form input raises the error flag (fr.Element is always canonical)
EC_HD Big8 msm_load_scalar(const MsmPlan& pl, const uint64_t* scalars, size_t i, uint32_t* err) {
    Big8 s = big_load(scalars + 4 * i);
    if (pl.scalars_mont) return f_from_mont<FrMod>(s);  // canonical whatever the input
    if (!f_is_canonical<FrMod>(s)) {
        ec_flag(err, MSM_ERR_SCALAR_RANGE);
        return big_zero();
    }
    return s;
}
// signed digit of window w: raw c bits + carry in; above 2^(c-1) it becomes raw - 2^c with a carry out
EC_HD int32_t msm_digit(const MsmPlan& pl, const uint32_t (&s)[9], uint32_t w, uint32_t& carry) {
    const uint32_t pos = w * pl.c, word = pos >> 5, off = pos & 31;
    uint32_t raw = 0;
    if (word < 8) {
        const uint64_t two = (uint64_t)s[word] | ((uint64_t)s[word + 1] << 32);  // s[8] == 0
        raw = (uint32_t)(two >> off) & ((1u << pl.c) - 1u);
    }
    raw += carry;
    if (raw > pl.B) {
        carry = 1;
        return (int32_t)raw - (int32_t)(1u << pl.c);
    }
    carry = 0;
    return (int32_t)raw;
}

// ---- kernel bodies ------------------------------------------------------------------------------------------------------------
struct KCount {  // i < n
    static EC_HD void run(size_t i, MsmPlan pl, const uint64_t* scalars, uint32_t* count, uint32_t* err) {
        const Big8 sc = msm_load_scalar(pl, scalars, i, err);
        uint32_t s[9];
#pragma unroll
        for (int k = 0; k < 8; k++) s[k] = sc.v[k];
        s[8] = 0;
        uint32_t carry = 0;
        for (uint32_t w = 0; w < pl.W; w++) {
            const int32_t d = msm_digit(pl, s, w, carry);
            if (d != 0) ec_atomic_inc(count + (size_t)w * pl.B + (uint32_t)(d < 0 ? -d : d) - 1u);
        }
        if (carry) ec_flag(err, MSM_ERR_INTERNAL);  // cannot happen for a canonical scalar (W * c >= 255)
    }
};
struct KScatter {  // i < n
    static EC_HD void run(size_t i, MsmPlan pl, const uint64_t* scalars, const uint32_t* off, uint32_t* cursor, uint32_t* entries, uint32_t* err) {
        const Big8 sc = msm_load_scalar(pl, scalars, i, err);
        uint32_t s[9];
#pragma unroll
        for (int k = 0; k < 8; k++) s[k] = sc.v[k];
        s[8] = 0;
        uint32_t carry = 0;
        for (uint32_t w = 0; w < pl.W; w++) {
            const int32_t d = msm_digit(pl, s, w, carry);
            if (d != 0) {
                const size_t key = (size_t)w * pl.B + (uint32_t)(d < 0 ? -d : d) - 1u;
                const uint32_t slot = ec_atomic_inc(cursor + key);
                entries[(size_t)off[key] + slot] = (uint32_t)i | (d < 0 ? 0x80000000u : 0u);
            }
        }
    }
};
struct KTaskCount {  // k < nkeys
    static EC_HD void run(size_t k, MsmPlan pl, const uint32_t* count, uint32_t* tcount) { tcount[k] = (count[k] + pl.T - 1) / pl.T; }
};
struct KGroupCount {  // k < nkeys: groups of at most G partial sums
    static EC_HD void run(size_t k, MsmPlan pl, const uint32_t* tcount, uint32_t* gcount) { gcount[k] = (tcount[k] + pl.G - 1) / pl.G; }
};
// exclusive prefix sum of in[0..n) into out[0..n], out[n] = total, in three steps over chunks of L elements
struct KScanA {  // j < nch
    static EC_HD void run(size_t j, const uint32_t* in, uint32_t* part, uint32_t n, uint32_t L) {
        const size_t lo = j * L, hi = lo + L < n ? lo + L : n;
        uint32_t s = 0;
        for (size_t e = lo; e < hi; e++) s += in[e];
        part[j] = s;
    }
};
struct KScanB {  // one thread
    static EC_HD void run(size_t, uint32_t* part, uint32_t nch) {
        uint32_t run = 0;
        for (uint32_t j = 0; j < nch; j++) {
            const uint32_t v = part[j];
            part[j] = run;
            run += v;
        }
        part[nch] = run;
    }
};
struct KScanC {  // j < nch
    static EC_HD void run(size_t j, const uint32_t* in, uint32_t* out, const uint32_t* part, uint32_t n, uint32_t L, uint32_t nch) {
        const size_t lo = j * L, hi = lo + L < n ? lo + L : n;
        uint32_t run = part[j];
        for (size_t e = lo; e < hi; e++) {
            out[e] = run;
            run += in[e];
        }
        if (j + 1 == nch) out[n] = part[nch];
    }
};
template <class C>
struct KAccum {  // t < max_tasks
    static EC_HD void run(size_t t, MsmPlan pl, const uint64_t* points, const uint32_t* entries, const uint32_t* off, const uint32_t* count,
                          const uint32_t* toff, uint64_t* partial) {
        typedef typename C::MInline M;
        if (t >= toff[pl.nkeys]) return;
        uint32_t lo = 0, hi = pl.nkeys;  // largest key with toff[key] <= t
        while (hi - lo > 1) {
            const uint32_t mid = lo + (hi - lo) / 2;
            if (toff[mid] <= t) lo = mid;
            else hi = mid;
        }
        const uint32_t key = lo;
        const size_t first = (size_t)off[key] + (t - toff[key]) * (size_t)pl.T;
        const size_t end_key = (size_t)off[key] + count[key];
        const size_t last = first + pl.T < end_key ? first + pl.T : end_key;
        typename C::X acc = C::x_inf();
        if (C::AFF_WORDS <= 8) {
            // the next point is requested before the current addition starts (one gather latency hidden per iteration)
            uint32_t e_cur = entries[first];
            typename C::Affine p_cur = C::aff_load(points + (size_t)C::AFF_WORDS * (e_cur & 0x7fffffffu));
            for (size_t e = first; e < last; e++) {
                uint32_t e_nxt = 0;
                typename C::Affine p_nxt = p_cur;
                if (e + 1 < last) {
                    e_nxt = entries[e + 1];
                    p_nxt = C::aff_load(points + (size_t)C::AFF_WORDS * (e_nxt & 0x7fffffffu));
                }
                if (e_cur & 0x80000000u) p_cur = C::aff_neg(p_cur);
                acc = C::template add_affine<M>(acc, p_cur);
                e_cur = e_nxt;
                p_cur = p_nxt;
            }
        } else {
            // G2: the point is 32 registers and the addition needs all the others; no prefetch
            for (size_t e = first; e < last; e++) {
                const uint32_t ee = entries[e];
                typename C::Affine p = C::aff_load(points + (size_t)C::AFF_WORDS * (ee & 0x7fffffffu));
                if (ee & 0x80000000u) p = C::aff_neg(p);
                acc = C::template add_affine<M>(acc, p);
            }
        }
        C::x_store(partial + (size_t)C::X_WORDS * t, acc);
    }
};
// Second level: a key with many tasks (a skewed input: thousands of partial sums for one bucket) is reduced in groups of G by
// one thread per GROUP, so that no single thread walks more than G partial sums here and (tasks of the key) / G group sums in KBucket.
template <class C>
struct KGroup {  // g < max_groups
    static EC_HD void run(size_t g, MsmPlan pl, const uint32_t* toff, const uint32_t* goff, const uint64_t* partial, uint64_t* gsum) {
        typedef typename C::MCall M;
        if (g >= goff[pl.nkeys]) return;
        uint32_t lo = 0, hi = pl.nkeys;  // largest key with goff[key] <= g
        while (hi - lo > 1) {
            const uint32_t mid = lo + (hi - lo) / 2;
            if (goff[mid] <= g) lo = mid;
            else hi = mid;
        }
        const uint32_t key = lo;
        const size_t first = (size_t)toff[key] + (g - goff[key]) * (size_t)pl.G;
        const size_t end_key = toff[key + 1];
        const size_t last = first + pl.G < end_key ? first + pl.G : end_key;
        typename C::X acc = C::x_load(partial + (size_t)C::X_WORDS * first);
        for (size_t t = first + 1; t < last; t++) acc = C::template add<M>(acc, C::x_load(partial + (size_t)C::X_WORDS * t));
        C::x_store(gsum + (size_t)C::X_WORDS * g, acc);
    }
};
template <class C>
struct KBucket {  // k < nkeys
    static EC_HD void run(size_t k, const uint32_t* goff, const uint64_t* gsum, uint64_t* bucket) {
        typedef typename C::MCall M;
        const uint32_t lo = goff[k], hi = goff[k + 1];
        typename C::X acc = C::x_inf();
        if (hi > lo) acc = C::x_load(gsum + (size_t)C::X_WORDS * lo);
        for (uint32_t t = lo + 1; t < hi; t++) acc = C::template add<M>(acc, C::x_load(gsum + (size_t)C::X_WORDS * t));
        C::x_store(bucket + (size_t)C::X_WORDS * k, acc);
    }
};
template <class C>
struct KChunk {  // j < W * nchunks
    static EC_HD void run(size_t j, MsmPlan pl, const uint64_t* bucket, uint64_t* chunk_out) {
        typedef typename C::MCall M;
        const uint32_t w = (uint32_t)(j / pl.nchunks), ci = (uint32_t)(j % pl.nchunks);
        const uint32_t lo = ci * pl.L, hi = lo + pl.L < pl.B ? lo + pl.L : pl.B;
        typename C::X run = C::x_inf(), acc = C::x_inf();
        for (uint32_t b = hi; b-- > lo;) {
            run = C::template add<M>(run, C::x_load(bucket + (size_t)C::X_WORDS * ((size_t)w * pl.B + b)));
            acc = C::template add<M>(acc, run);
        }
        // acc = sum (b - lo + 1) S_b; bucket b weighs b + 1
        if (lo) acc = C::template add<M>(acc, C::template mul_small<M>(run, lo));
        C::x_store(chunk_out + (size_t)C::X_WORDS * j, acc);
    }
};
template <class C>
struct KWindow {  // w < W
    static EC_HD void run(size_t w, MsmPlan pl, const uint64_t* chunk_out, uint64_t* win) {
        typedef typename C::MCall M;
        typename C::X acc = C::x_inf();
        for (uint32_t ci = 0; ci < pl.nchunks; ci++) acc = C::template add<M>(acc, C::x_load(chunk_out + (size_t)C::X_WORDS * (w * pl.nchunks + ci)));
        C::x_store(win + (size_t)C::X_WORDS * w, acc);
    }
};
// out[0 .. AFF_WORDS): affine result, Montgomery (the G1Affine / G2Affine memory image); out[AFF_WORDS .. 2 AFF_WORDS): the same in
// regular form (for RawBytes)
template <class C>
EC_HD void msm_store_result(uint64_t* out, const typename C::X& r) {
    const typename C::Affine a = C::to_affine(r);
    C::aff_store(out, a);
    typename C::Affine reg;
    reg.x = C::Field::from_mont(a.x), reg.y = C::Field::from_mont(a.y);
    C::aff_store(out + C::AFF_WORDS, reg);
}
template <class C>
struct KFinal {  // one thread
    static EC_HD void run(size_t, MsmPlan pl, const uint64_t* win, uint64_t* out) {
        typedef typename C::MCall M;
        typename C::X acc = C::x_inf();
        for (uint32_t w = pl.W; w-- > 0;) {
            for (uint32_t k = 0; k < pl.c; k++) acc = C::template dbl<M>(acc);
            acc = C::template add<M>(acc, C::x_load(win + (size_t)C::X_WORDS * w));
        }
        msm_store_result<C>(out, acc);
    }
};
template <class C>
struct KAddAffine {  // one thread: out = a + b (G1Affine.Add, hints.go:184), same output format as KFinal
    static EC_HD void run(size_t, const uint64_t* a, const uint64_t* b, uint64_t* out) {
        typedef typename C::MCall M;
        const typename C::X s = C::template add_affine<M>(C::from_affine(C::aff_load(a)), C::aff_load(b));
        msm_store_result<C>(out, s);
    }
};

// ---- workspace and driver -----------------------------------------------------------------------------------------------------
struct MsmWorkspace {  // offsets into one device buffer
    size_t count, off, cursor, tcount, toff, gcount, goff, part, entries, partial, gsum, bucket, chunk_out, win, err, out, bytes;
};
inline size_t msm_align(size_t x) { return (x + 255) & ~(size_t)255; }
inline MsmWorkspace msm_layout(const MsmPlan& pl, size_t x_bytes, size_t aff_bytes) {  // sizes of one XYZZ / affine image
    MsmWorkspace ws;
    size_t o = 0;
    const size_t nk = pl.nkeys;
    const size_t nch = (nk + pl.scan_L - 1) / pl.scan_L;
    ws.count = o, o = msm_align(o + 4 * (nk + 1));
    ws.cursor = o, o = msm_align(o + 4 * nk);
    ws.err = o, o = msm_align(o + 4);  // count, cursor, err are contiguous: one memset
    ws.off = o, o = msm_align(o + 4 * (nk + 1));
    ws.tcount = o, o = msm_align(o + 4 * (nk + 1));
    ws.toff = o, o = msm_align(o + 4 * (nk + 1));
    ws.gcount = o, o = msm_align(o + 4 * (nk + 1));
    ws.goff = o, o = msm_align(o + 4 * (nk + 1));
    ws.part = o, o = msm_align(o + 4 * (nch + 1));
    ws.entries = o, o = msm_align(o + 4 * ((size_t)pl.n * pl.W + 1));
    ws.partial = o, o = msm_align(o + x_bytes * pl.max_tasks);
    ws.gsum = o, o = msm_align(o + x_bytes * pl.max_groups);
    ws.bucket = o, o = msm_align(o + x_bytes * nk);
    ws.chunk_out = o, o = msm_align(o + x_bytes * (size_t)pl.W * pl.nchunks);
    ws.win = o, o = msm_align(o + x_bytes * pl.W);
    ws.out = o, o = msm_align(o + 2 * aff_bytes);
    ws.bytes = o;
    return ws;
}

// Enqueues one multi-exponentiation on curve C (G1 or G2) on the executor.  `base` is a device buffer of at least
// msm_layout(pl, 8 * C::X_WORDS, 8 * C::AFF_WORDS).bytes; the result (2 * AFF_WORDS words, see msm_store_result) lands at
// base + ws.out, the error flag at base + ws.err.  Returns the number of launches.
template <class C, class Exec>
int msm_enqueue(Exec& ex, const MsmPlan& pl, const MsmWorkspace& ws, unsigned char* base, const uint64_t* d_points, const uint64_t* d_scalars) {
    uint32_t* count = (uint32_t*)(base + ws.count);
    uint32_t* cursor = (uint32_t*)(base + ws.cursor);
    uint32_t* err = (uint32_t*)(base + ws.err);
    uint32_t* off = (uint32_t*)(base + ws.off);
    uint32_t* tcount = (uint32_t*)(base + ws.tcount);
    uint32_t* toff = (uint32_t*)(base + ws.toff);
    uint32_t* gcount = (uint32_t*)(base + ws.gcount);
    uint32_t* goff = (uint32_t*)(base + ws.goff);
    uint32_t* part = (uint32_t*)(base + ws.part);
    uint32_t* entries = (uint32_t*)(base + ws.entries);
    uint64_t* partial = (uint64_t*)(base + ws.partial);
    uint64_t* gsum = (uint64_t*)(base + ws.gsum);
    uint64_t* bucket = (uint64_t*)(base + ws.bucket);
    uint64_t* chunk_out = (uint64_t*)(base + ws.chunk_out);
    uint64_t* win = (uint64_t*)(base + ws.win);
    uint64_t* out = (uint64_t*)(base + ws.out);
    const uint32_t nk = pl.nkeys, sl = pl.scan_L, nch = (nk + sl - 1) / sl;
    int launches = 0;
    ex.zero(base + ws.count, ws.off - ws.count);  // count, cursor, err
    launches += ex.template launch<KCount>(pl.n, pl, d_scalars, count, err);
    launches += ex.template launch<KScanA>(nch, (const uint32_t*)count, part, nk, sl);
    launches += ex.template launch<KScanB>(1, part, nch);
    launches += ex.template launch<KScanC>(nch, (const uint32_t*)count, off, (const uint32_t*)part, nk, sl, nch);
    launches += ex.template launch<KScatter>(pl.n, pl, d_scalars, (const uint32_t*)off, cursor, entries, err);
    launches += ex.template launch<KTaskCount>(nk, pl, (const uint32_t*)count, tcount);
    launches += ex.template launch<KScanA>(nch, (const uint32_t*)tcount, part, nk, sl);
    launches += ex.template launch<KScanB>(1, part, nch);
    launches += ex.template launch<KScanC>(nch, (const uint32_t*)tcount, toff, (const uint32_t*)part, nk, sl, nch);
    launches += ex.template launch<KAccum<C>>(pl.max_tasks, pl, d_points, (const uint32_t*)entries, (const uint32_t*)off, (const uint32_t*)count,
                                              (const uint32_t*)toff, partial);
    launches += ex.template launch<KGroupCount>(nk, pl, (const uint32_t*)tcount, gcount);
    launches += ex.template launch<KScanA>(nch, (const uint32_t*)gcount, part, nk, sl);
    launches += ex.template launch<KScanB>(1, part, nch);
    launches += ex.template launch<KScanC>(nch, (const uint32_t*)gcount, goff, (const uint32_t*)part, nk, sl, nch);
    launches += ex.template launch<KGroup<C>>(pl.max_groups, pl, (const uint32_t*)toff, (const uint32_t*)goff, (const uint64_t*)partial, gsum);
    launches += ex.template launch<KBucket<C>>(nk, (const uint32_t*)goff, (const uint64_t*)gsum, bucket);
    launches += ex.template launch<KChunk<C>>((size_t)pl.W * pl.nchunks, pl, (const uint64_t*)bucket, chunk_out);
    launches += ex.template launch<KWindow<C>>(pl.W, pl, (const uint64_t*)chunk_out, win);
    launches += ex.template launch<KFinal<C>>(1, pl, (const uint64_t*)win, out);
    return launches;
}

}  // namespace ec
