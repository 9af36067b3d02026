// Device kernels of the GKR-MiMC prover (sm_100a).  One kernel per CPU hot loop of the reference:
//   K1 k_mimc_assign   <- Circuit.Assign / Layer.Evaluate / CipherGate.EvalBatch   (circuit/assignment.go:12-32,
//                         circuit/circuit.go:48-64, circuit/gates/cipher.go:25-42)
//   K2 k_eq_suffix / k_eq_small + k_eq_expand <- poly.FoldedEqTable / ChunkOfEqTable (poly/eq.go:41-89)
//   K5 k_eq_expand (n_claims>1) <- multi-claim combination                         (sumcheck/prover.go:121-141, algo.go:219-223)
//   K3 k_round_cf (factored cipher round) / k_round (generic) <- getPartialPolyChunk + consumeAccumulate
//                                                                                  (sumcheck/algo.go:54-205, prover.go:236-245)
//   K4 fold fused into the round kernels / k_fold <- MultiLin.FoldChunk            (poly/multilin.go:26-36, sumcheck/algo.go:46-51)
//   staging: k_destripe, k_take_shard (multi-GPU input shards), k_convert, k_hash_out (hint I/O, prover/gadget/hints.go:197-233)
// All tables are arrays of FrRaw (Go's fr.Element image), moved with 256-bit accesses.  Results leave the device as tagged 64-bit
// words (publish_word); the verifier challenge can come back the same way (ChalWait).
#pragma once
#include <cuda_runtime.h>

#include "fr_device.cuh"

namespace gkr {

enum { GATE_IDENTITY = 0, GATE_CIPHER = 1 };

__constant__ FrRaw c_arks[91] = {
#include "mimc_arks.inc"
};

// ------------------------------------------------------------------------------------------------
// Block-wide / grid-wide reduction of NACC field accumulators per thread.
// Shared layout u32 sm[NACC][8][BLOCK] (thread index fastest => conflict-free).  Field addition is
// exact, so the summation order is irrelevant to the result (no determinism concern).
// ------------------------------------------------------------------------------------------------
template <int BLOCK>
__device__ __forceinline__ Fr smem_get(const uint32_t* sm, int k, int tid) {
    Fr a;
#pragma unroll
    for (int l = 0; l < 8; l++) a.v[l] = sm[(k * 8 + l) * BLOCK + tid];
    return a;
}
// tree-reduce sm[k][.][0..BLOCK) into sm[k][.][0] for every k, using all threads at every level
template <int NACC, int BLOCK>
__device__ __forceinline__ void smem_tree_reduce(uint32_t* sm, int tid) {
    __syncthreads();
#pragma unroll 1
    for (int stride = BLOCK / 2; stride >= 1; stride >>= 1) {
        const int items = NACC * stride;
#pragma unroll 1
        for (int it = tid; it < items; it += BLOCK) {
            const int k = it / stride, i = it - k * stride;
            Fr a = smem_get<BLOCK>(sm, k, i), b = smem_get<BLOCK>(sm, k, i + stride);
            a = fr_add(a, b);
#pragma unroll
            for (int l = 0; l < 8; l++) sm[(k * 8 + l) * BLOCK + i] = a.v[l];
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// Publishing results and receiving the verifier challenge.
//
// Publishing without fences: every 64-bit word carries the launch tag in its upper half and is written with one aligned
// 8-byte store, so the consumer (a host thread spinning on mapped pinned memory, or a later kernel) needs no flag and no
// ordering between words -- it waits until all words show the tag.  Saves the two system-scope fences (~3 us measured) a
// flag protocol costs per round.  The challenge r_k = MimcHash(round polynomial) comes back the same way: the host that
// owns the proof's transcript writes 8 tagged words into mapped memory and the LAST block of the round kernel (all other
// blocks have exited, the SMs are free for other streams) polls them and stores r_k to device memory, where the next
// launch on the stream -- already enqueued -- reads it.  No host thread of this GPU takes part in a round.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void publish_word(unsigned long long* dst, uint32_t seq, uint32_t limb) {
    const unsigned long long v = ((unsigned long long)seq << 32) | limb;
    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(dst), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long poll_word(const unsigned long long* src) {
    unsigned long long v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(src) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
struct ChalWait {
    const unsigned long long* src;  // 8 tagged words (mapped host memory, device alias); nullptr: the kernel does not wait
    FrRaw* dst;                     // device memory that receives the challenge (read by the next launch on the stream)
    unsigned int* err;              // device flag: set to 1 when the challenge did not arrive within CHAL_TIMEOUT_NS
    uint32_t tag;
};
constexpr unsigned long long CHAL_TIMEOUT_NS = 20ull * 1000 * 1000 * 1000;  // never hang the box: give up after 20 s
// called by every thread of the one block that published; lanes 0..7 each fetch one limb
__device__ __forceinline__ void wait_challenge(const ChalWait& w) {
    if (!w.src || threadIdx.x >= 8) return;
    const unsigned long long t0 = global_timer_ns();
    unsigned long long v = poll_word(w.src + threadIdx.x);
    while ((uint32_t)(v >> 32) != w.tag) {
        __nanosleep(200);
        v = poll_word(w.src + threadIdx.x);
        if (global_timer_ns() - t0 > CHAL_TIMEOUT_NS) {
            atomicExch(w.err, 1u);
            break;
        }
    }
    reinterpret_cast<uint32_t*>(w.dst)[threadIdx.x] = (uint32_t)v;
}

// Output of the generic round kernel: NEV canonical field elements, each as 9 tagged words (8 limbs + a zero word: the same
// 288-bit image the factored kernel's reduced sums use, so one host routine reads both).
struct ReduceOut {
    FrRaw* partials;             // [gridDim.x][NEV] device scratch
    unsigned int* ticket;        // device counter, zero between launches
    unsigned long long* result;  // NEV x 9 tagged words; device memory or mapped host memory
    uint32_t seq;
    ChalWait chal;
};

// ------------------------------------------------------------------------------------------------
// K1: batched MiMC layer assignment.  One thread = one hash; the 91 round states go to layers 3..93.
// layers[l] for l in 3..93 is at base + (l-3)*stride.  a[2] is an alias of a[0] (IdentityGate copy).
// Per hash: 364 Fr-mul, reads 64 B, writes 91*32 B.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_mimc_assign(const FrRaw* __restrict__ key, const FrRaw* __restrict__ msg, FrRaw* __restrict__ base,
                                                      size_t stride, size_t n) {
    for (size_t x = (size_t)blockIdx.x * blockDim.x + threadIdx.x; x < n; x += (size_t)gridDim.x * blockDim.x) {
        const Fr k = fr_load_stream(key + x);
        Fr s = fr_load_stream(msg + x);
#pragma unroll 1
        for (int i = 0; i < 91; i++) {
            // examples/mimc.go:32: layer i+3 = Cipher(Arks[i])(a[2], prev) = (prev + ark + key)^7 (cipher.go:34-40)
            Fr t = fr_add(fr_add(s, fr_unpack(c_arks[i])), k);
            s = fr_pow7(t);
            fr_store(base + (size_t)i * stride + x, s);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K2/K5 stage 1: small eq tables by doubling (poly/eq.go:49-56), one block per (claim, half).
// For claim j: hi_j = mult_j * eq(q_j[0:nh], .)  (2^nh entries),  lo_j = eq(q_j[nh:nh+nl], .)  (2^nl entries).
// blockIdx.x = 2*j + (0: hi, 1: lo).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_eq_small(const FrRaw* __restrict__ qprimes, int bn, int nh, int nl, const FrRaw* __restrict__ mults,
                                                   FrRaw* __restrict__ hi, FrRaw* __restrict__ lo) {
    const int j = blockIdx.x >> 1, which = blockIdx.x & 1;
    const FrRaw* q = qprimes + (size_t)j * bn + (which ? nh : 0);
    const int nv = which ? nl : nh;
    FrRaw* t = which ? lo + ((size_t)j << nl) : hi + ((size_t)j << nh);
    if (threadIdx.x == 0) fr_store(t, (which == 0 && mults) ? fr_load(mults + j) : fr_one());
    __syncthreads();
    for (int i = 0; i < nv; i++) {
        const Fr r = fr_load(q + i);
        for (int jj = threadIdx.x; jj < (1 << i); jj += blockDim.x) {
            const size_t J = (size_t)jj << (nv - i);
            const size_t JN = J + ((size_t)1 << (nv - 1 - i));
            const Fr a = fr_load(t + J);
            const Fr b = fr_mul(r, a);
            fr_store(t + JN, b);
            fr_store(t + J, fr_sub(a, b));
        }
        __syncthreads();
    }
}
// K2/K5 stage 2: out[x] = sum_j hi_j[x >> nl] * lo_j[x & (2^nl-1)]   (one streaming write of the table)
// The 91 products of an entry only feed a sum, so they are taken as PLAIN 512-bit products (64 instead of 136 wide
// multiply-adds each, fr_mul_acc_wide on a 17-limb accumulator in shared memory) and reduced in groups of five: 5*q^2 <
// q*2^256 is exactly the input range of one Montgomery reduction (fr_redc_wide), and REDC(sum of products) is the sum of the
// Montgomery products, so the result is the same canonical element.  91 claims: 7 192 instead of 12 376 wide multiply-adds.
constexpr int EQX_BLOCK = 256, EQX_GROUP = 5;
__global__ void __launch_bounds__(EQX_BLOCK) k_eq_expand(const FrRaw* __restrict__ hi, const FrRaw* __restrict__ lo, int nh, int nl, int n_claims,
                                                         FrRaw* __restrict__ out, size_t n) {
    __shared__ uint32_t acc[17 * EQX_BLOCK];
    const size_t mask = ((size_t)1 << nl) - 1;
    const int tid = threadIdx.x;
    if (n_claims == 1) {  // poly.FoldedEqTable proper: one product per entry, one streaming write
        for (size_t x = (size_t)blockIdx.x * blockDim.x + tid; x < n; x += (size_t)gridDim.x * blockDim.x)
            fr_store(out + x, fr_mul(fr_load(hi + (x >> nl)), fr_load(lo + (x & mask))));
        return;
    }
    for (size_t x = (size_t)blockIdx.x * blockDim.x + tid; x < n; x += (size_t)gridDim.x * blockDim.x) {
        Fr total = fr_zero();
#pragma unroll 1
        for (int j0 = 0; j0 < n_claims; j0 += EQX_GROUP) {
#pragma unroll
            for (int l = 0; l < 17; l++) acc[l * EQX_BLOCK + tid] = 0;
            const int j1 = min(j0 + EQX_GROUP, n_claims);
#pragma unroll 1
            for (int j = j0; j < j1; j++) {
                const Fr h = fr_load(hi + ((size_t)j << nh) + (x >> nl));
                const Fr l = fr_load(lo + ((size_t)j << nl) + (x & mask));
                fr_mul_acc_wide(acc + tid, EQX_BLOCK, h, l);
            }
            uint32_t t[16];
#pragma unroll
            for (int l = 0; l < 16; l++) t[l] = acc[l * EQX_BLOCK + tid];  // limb 16 stays zero: 5*q^2 < 2^512
            total = fr_add(total, fr_redc_wide(t));
        }
        fr_store(out + x, total);
    }
}

// ------------------------------------------------------------------------------------------------
// K4 standalone: fold NT tables: dst[i] = src[i] + r*(src[i+half] - src[i]),  i < half.
// ------------------------------------------------------------------------------------------------
struct FoldArgs {
    const FrRaw* src[3];
    FrRaw* dst[3];
    int n_tables;
    size_t half;
    FrRaw r;
    const FrRaw* r_dev;  // when not null: the challenge is read from device memory instead of a.r
    FrConstMul rk;       // r_dev == nullptr: the constant-multiplier table of r (fr_mul_const)
};
__global__ void __launch_bounds__(256) k_fold(const FoldArgs a) {
    const Fr r = a.r_dev ? fr_load(a.r_dev) : fr_unpack(a.r);
    const size_t total = a.half * (size_t)a.n_tables;
    const size_t step = (size_t)gridDim.x * blockDim.x;
    // two outputs per iteration, all four loads in flight before the first product: at 96 B per 136 wide multiply-adds the kernel
    // sits at B200's ridge (6.5 TB/s / 8.1 T wide MAC/s = 0.8 B per MAC), so neither DRAM latency nor the multiplier may idle
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += 2 * step) {
        const size_t i2 = i + step;
        const bool two = i2 < total;
        const int t = (int)(i / a.half), t2 = two ? (int)(i2 / a.half) : t;
        const size_t x = i - (size_t)t * a.half, x2 = two ? i2 - (size_t)t2 * a.half : x;
        // table pointers by selection, not by indexing the parameter arrays (that would copy them to local memory)
        const FrRaw* s1 = t == 0 ? a.src[0] : (t == 1 ? a.src[1] : a.src[2]);
        const FrRaw* s2 = t2 == 0 ? a.src[0] : (t2 == 1 ? a.src[1] : a.src[2]);
        FrRaw* d1 = t == 0 ? a.dst[0] : (t == 1 ? a.dst[1] : a.dst[2]);
        FrRaw* d2 = t2 == 0 ? a.dst[0] : (t2 == 1 ? a.dst[1] : a.dst[2]);
        const Fr b = fr_load_stream(s1 + x), u = fr_load_stream(s1 + x + a.half);
        const Fr b2 = fr_load_stream(s2 + x2), u2 = fr_load_stream(s2 + x2 + a.half);
        if (a.r_dev) {
            fr_store(d1 + x, fr_add(b, fr_mul(r, fr_sub(u, b))));
            if (two) fr_store(d2 + x2, fr_add(b2, fr_mul(r, fr_sub(u2, b2))));
        } else {
            fr_store(d1 + x, fr_add(b, fr_mul_const(a.rk, fr_sub(u, b))));
            if (two) fr_store(d2 + x2, fr_add(b2, fr_mul_const(a.rk, fr_sub(u2, b2))));
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K3(+K4): one sumcheck round.  FOLD: first fold the source tables (length 4*half) with the previous
// challenge r into dst (length 2*half; in place allowed), then evaluate this round's polynomial on the
// `half` pairs (x, x+half) of the folded tables.  !FOLD: evaluate directly on src (length 2*half).
//   evals[t] = sum_x eq_t(x) * gate(X_t(x)),  v_t(x) = v[x] + t*(v[x+half]-v[x]),  t = 0..NEV-1
// cipher: gate = (X0 + X1 + ark)^7, NEV = 9 (45 mul/pair + 6 for the fold); identity: gate = X0, NEV = 3.
// ------------------------------------------------------------------------------------------------
struct RoundArgs {
    const FrRaw* src[3];  // eq, X0, X1
    FrRaw* dst[3];
    size_t half;
    FrRaw r;             // previous challenge (FOLD) ...
    const FrRaw* r_dev;  // ... or, when not null, where a previous launch left it in device memory
    FrRaw ark;           // cipher gate constant
    ReduceOut red;
    uint32_t* partials_w;  // identity gate: 64-bit limb-column sums of the wide (unreduced) round sums, zero between launches
};

template <int GATE>
struct GateTraits;
template <>
struct GateTraits<GATE_IDENTITY> {
    static constexpr int NIN = 1, NEV = 3;
};
template <>
struct GateTraits<GATE_CIPHER> {
    static constexpr int NIN = 2, NEV = 9;
};

template <bool FOLD>
__device__ __forceinline__ void load_pair(const FrRaw* src, FrRaw* dst, size_t x, size_t half, const Fr& r, Fr& bot, Fr& top) {
    if (FOLD) {
        const size_t m = 2 * half;
        const Fr b0 = fr_load_stream(src + x), t0 = fr_load_stream(src + x + m);
        const Fr b1 = fr_load_stream(src + x + half), t1 = fr_load_stream(src + x + half + m);
        bot = fr_add(b0, fr_mul(r, fr_sub(t0, b0)));
        top = fr_add(b1, fr_mul(r, fr_sub(t1, b1)));
        fr_store(dst + x, bot);
        fr_store(dst + x + half, top);
    } else {
        bot = fr_load_stream(src + x);
        top = fr_load_stream(src + x + half);
    }
}

// ---- lazy 288-bit accumulators in shared memory -------------------------------------------------
// Each thread owns NEV accumulators of 9 x 32-bit limbs: sm[(t*9 + l)*BLOCK + tid].  Terms (< q < 2^254) are
// added as plain integers without reduction (a grid holds < 2^24 terms per accumulator, so 288 bits never
// overflow); one reduction mod q per block at the end.  Keeps the register file free for the multiplier.
template <int BLOCK>
__device__ __forceinline__ void wide_acc_add(uint32_t* slot /* &sm[t*9*BLOCK + tid] */, const Fr& term) {
    uint32_t a[9];
#pragma unroll
    for (int l = 0; l < 9; l++) a[l] = slot[l * BLOCK];
    asm("add.cc.u32 %0, %0, %9;\n\t"
        "addc.cc.u32 %1, %1, %10;\n\t"
        "addc.cc.u32 %2, %2, %11;\n\t"
        "addc.cc.u32 %3, %3, %12;\n\t"
        "addc.cc.u32 %4, %4, %13;\n\t"
        "addc.cc.u32 %5, %5, %14;\n\t"
        "addc.cc.u32 %6, %6, %15;\n\t"
        "addc.cc.u32 %7, %7, %16;\n\t"
        "addc.u32 %8, %8, 0;"
        : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]), "+r"(a[7]), "+r"(a[8])
        : "r"(term.v[0]), "r"(term.v[1]), "r"(term.v[2]), "r"(term.v[3]), "r"(term.v[4]), "r"(term.v[5]), "r"(term.v[6]), "r"(term.v[7]));
#pragma unroll
    for (int l = 0; l < 9; l++) slot[l * BLOCK] = a[l];
}
// sm[k][.][i] += sm[k][.][i + stride]  as 288-bit integers
template <int BLOCK>
__device__ __forceinline__ void wide_pair_add(uint32_t* sm, int k, int i, int stride) {
    uint32_t a[9], b[9];
#pragma unroll
    for (int l = 0; l < 9; l++) {
        a[l] = sm[(k * 9 + l) * BLOCK + i];
        b[l] = sm[(k * 9 + l) * BLOCK + i + stride];
    }
    asm("add.cc.u32 %0, %0, %9;\n\t"
        "addc.cc.u32 %1, %1, %10;\n\t"
        "addc.cc.u32 %2, %2, %11;\n\t"
        "addc.cc.u32 %3, %3, %12;\n\t"
        "addc.cc.u32 %4, %4, %13;\n\t"
        "addc.cc.u32 %5, %5, %14;\n\t"
        "addc.cc.u32 %6, %6, %15;\n\t"
        "addc.cc.u32 %7, %7, %16;\n\t"
        "addc.u32 %8, %8, %17;"
        : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]), "+r"(a[7]), "+r"(a[8])
        : "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]), "r"(b[8]));
#pragma unroll
    for (int l = 0; l < 9; l++) sm[(k * 9 + l) * BLOCK + i] = a[l];
}
// 288-bit integer (9 limbs) -> canonical residue:  lo + hi*2^256  mod q
__device__ __forceinline__ Fr fr_from_wide(const uint32_t (&w)[9]) {
    Fr lo, hi = fr_zero(), r2, raw1 = fr_zero();
#pragma unroll
    for (int l = 0; l < 8; l++) lo.v[l] = w[l];
    hi.v[0] = w[8];
    raw1.v[0] = 1;
    // R^2 mod q
    r2.v[0] = 0xae216da7u; r2.v[1] = 0x1bb8e645u; r2.v[2] = 0xe35c59e3u; r2.v[3] = 0x53fe3ab1u;
    r2.v[4] = 0x53bb8085u; r2.v[5] = 0x8c49833du; r2.v[6] = 0x7f4e44a5u; r2.v[7] = 0x0216d0b1u;
    // Montgomery products with one factor < q are valid for any 256-bit other factor (result < 2q, reduced once)
    const Fr lo_red = fr_mul(fr_mul(lo, r2), raw1);  // lo*R*R^-1 = lo mod q
    const Fr hi_red = fr_mul(hi, r2);                // hi*2^256 mod q
    return fr_add(lo_red, hi_red);
}

template <int NEV, int BLOCK>
__device__ __forceinline__ void grid_reduce_wide(uint32_t* sm, const ReduceOut& out) {
    const int tid = threadIdx.x;
    __syncthreads();
    // block tree over threads, all NEV accumulators at every level
#pragma unroll 1
    for (int stride = BLOCK / 2; stride >= 1; stride >>= 1) {
        const int items = NEV * stride;
#pragma unroll 1
        for (int it = tid; it < items; it += BLOCK) {
            const int k = it / stride, i = it - k * stride;
            wide_pair_add<BLOCK>(sm, k, i, stride);
        }
        __syncthreads();
    }
    __shared__ bool is_last_w;
    Fr mine = fr_zero();
    if (tid < NEV) {
        uint32_t w[9];
#pragma unroll
        for (int l = 0; l < 9; l++) w[l] = sm[(tid * 9 + l) * BLOCK];
        mine = fr_from_wide(w);
    }
    if (gridDim.x > 1) {
        if (tid < NEV) fr_store(out.partials + (size_t)blockIdx.x * NEV + tid, mine);
        __threadfence();
        __syncthreads();
        if (tid == 0) is_last_w = (atomicAdd(out.ticket, 1u) == gridDim.x - 1);
        __syncthreads();
        if (!is_last_w) return;
        __threadfence();
        // last block: modular sum of the per-block partials (reuses the shared buffer with the Fr layout)
#pragma unroll 1
        for (int k = 0; k < NEV; k++) {
            Fr s = fr_zero();
#pragma unroll 1
            for (unsigned b = tid; b < gridDim.x; b += BLOCK) s = fr_add(s, fr_load(out.partials + (size_t)b * NEV + k));
#pragma unroll
            for (int l = 0; l < 8; l++) sm[(k * 8 + l) * BLOCK + tid] = s.v[l];
        }
        smem_tree_reduce<NEV, BLOCK>(sm, tid);
        if (tid < NEV) mine = smem_get<BLOCK>(sm, tid, 0);
        if (tid == 0) *out.ticket = 0;
    }
    if (tid < NEV) {
#pragma unroll
        for (int l = 0; l < 8; l++) publish_word(out.result + tid * 9 + l, out.seq, mine.v[l]);
        publish_word(out.result + tid * 9 + 8, out.seq, 0u);
    }
    wait_challenge(out.chal);
}

// ================================================================================================
// Factored cipher round (K3+K4, single-claim layers: 91 of the 92 sumchecks of the MiMC circuit).
//
// For one claim the eq table is a product, eq(q,x) = prod_j eq(q_j,x_j), so in round k (variables 0..k-1
// already bound to r_0..r_{k-1}, MSB first as in poly/multilin.go:26-36) the folded table of
// sumcheck/prover.go is   Eq_k[x_k, x'] = c_k * eq(q_k, x_k) * T_k[x'],   c_k = prod_{j<k} eq(q_j, r_j),
// T_k = eq(q[k+1:], .).  The round polynomial of sumcheck/algo.go:54-205 therefore factors as
//      p_k(t) = c_k * ((1-q_k)(1-t) + q_k t) * S_k(t),     S_k(t) = sum_x' T_k[x'] * (a(x') + t*b(x'))^7,
// with a = X0[x'] + X1[x'] + ark (bottom half) and b = (top half) - a.  The device only produces the
// coefficient sums  m_i = sum_x' T_k[x'] * a^(7-i) * b^i  (S_k(t) = sum_i C(7,i) m_i t^i): 18 field
// multiplications per pair instead of the 45 of the evaluate-at-9-points form, no eq table in HBM, no eq
// fold.  m_7 follows on the host from the running claim (p_k(0)+p_k(1) = claim); NM = 8 computes it on the
// device for the rounds where no claim is available.  The host turns m into the very coefficients the
// reference sends (exact field arithmetic: identical canonical values), see Ctx::sumcheck.
//
// T_k[x'] = A[x' >> c] * B[x' & (2^c-1)]  from two small suffix tables built per layer by k_eq_suffix
// (or T_k = B[x'] once x' has <= c bits).  PAR = 8 spreads one pair over 8 lanes (lane j forms term j
// with a uniform select-and-multiply ladder) so that small rounds are not bound by one thread's chain
// of ~23 dependent multiplications.
// ================================================================================================
#ifdef GKR_TRACE
__device__ long long g_trace[16];
#define GKR_T(i) do { if (blockIdx.x == 0 && threadIdx.x == 0) g_trace[i] = clock64(); } while (0)
#else
#define GKR_T(i) do { } while (0)
#endif

// a += b as WL-limb integers (WL = 9 or 17); every carry chain is one asm statement, carries cross statements in a register
template <int WL>
__device__ __forceinline__ void widen_add(uint32_t (&a)[WL], const uint32_t (&b)[WL]) {
    static_assert(WL == 9 || WL == 17, "288- or 544-bit accumulators");
    uint32_t c = add8_carry(a, b);
    if (WL == 17) c = add8_carry_in(a + 8, b + 8, c);
    a[WL - 1] += b[WL - 1] + c;
}
// sm[(k*WL + l)*BLOCK + i] += sm[(k*WL + l)*BLOCK + i + stride]  as WL-limb integers
template <int WL, int BLOCK>
__device__ __forceinline__ void widen_pair_add(uint32_t* sm, int k, int i, int stride) {
    uint32_t a[WL], b[WL];
#pragma unroll
    for (int l = 0; l < WL; l++) {
        a[l] = sm[(k * WL + l) * BLOCK + i];
        b[l] = sm[(k * WL + l) * BLOCK + i + stride];
    }
    widen_add<WL>(a, b);
#pragma unroll
    for (int l = 0; l < WL; l++) sm[(k * WL + l) * BLOCK + i] = a[l];
}

struct WideOut {
    uint32_t* partials;            // NM x WL 64-bit column sums (device scratch, zero between launches)
    unsigned int* ticket;          // device counter, zero between launches
    unsigned long long* result;    // NM x 9 words, each (seq << 32) | limb of a 288-bit plain sum (NOT reduced mod q);
                                   // device memory or mapped host memory
    uint32_t seq;                  // tag of this launch: the consumer polls until every word carries it
    ChalWait chal;                 // optional: the block that publishes then waits for the challenge derived from the sums
};

// a += b as 288-bit integers (9 x 32-bit limbs)
__device__ __forceinline__ void wide9_add(uint32_t (&a)[9], const uint32_t (&b)[9]) {
    asm("add.cc.u32 %0, %0, %9;\n\t"
        "addc.cc.u32 %1, %1, %10;\n\t"
        "addc.cc.u32 %2, %2, %11;\n\t"
        "addc.cc.u32 %3, %3, %12;\n\t"
        "addc.cc.u32 %4, %4, %13;\n\t"
        "addc.cc.u32 %5, %5, %14;\n\t"
        "addc.cc.u32 %6, %6, %15;\n\t"
        "addc.cc.u32 %7, %7, %16;\n\t"
        "addc.u32 %8, %8, %17;"
        : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]), "+r"(a[7]), "+r"(a[8])
        : "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]), "r"(b[8]));
}

// Grid stage shared by both layouts of k_round_cf.  tot: this block's NM x 9 limb sums in shared memory
// (tot[k*WL + l]); scratch: >= 2 * NM * WL words of shared memory.  No field multiplication anywhere:
// the host reduces the NM wide sums modulo q.
template <int NM, int WL, int BLOCK>
__device__ __forceinline__ void grid_stage_wide(const uint32_t* tot, uint32_t* scratch, const WideOut& out) {
    const int tid = threadIdx.x;
    __shared__ bool is_last_r;
    GKR_T(5);
    if (gridDim.x == 1) {
        for (int i = tid; i < NM * WL; i += BLOCK) publish_word(out.result + i, out.seq, tot[i]);
        GKR_T(6);
        wait_challenge(out.chal);
        return;
    }
    // Every block adds its NM x WL limbs to 64-bit COLUMN sums in global memory (RED.ADD.64, fire and forget: a column
    // holds < gridDim.x * 2^32, no carries between columns yet); the last block to arrive propagates the carries once.
    // The earlier scheme (per-block partials summed by the last block) cost ~20 us per 444-block launch: one SM
    // walking 200 KB of partials at L2 latency.
    unsigned long long* gacc = reinterpret_cast<unsigned long long*>(out.partials);
    for (int i = tid; i < NM * WL; i += BLOCK) atomicAdd(gacc + i, (unsigned long long)tot[i]);
    __threadfence();
    __syncthreads();
    if (tid == 0) is_last_r = (atomicAdd(out.ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!is_last_r) return;
    __threadfence();
    for (int i = tid; i < NM * WL; i += BLOCK) {
        const unsigned long long v = __ldcg(gacc + i);
        gacc[i] = 0;  // ready for the next launch on this stream
        scratch[i] = (uint32_t)v;
        scratch[NM * WL + i] = (uint32_t)(v >> 32);
    }
    __syncthreads();
    if (tid < NM) {
        unsigned long long carry = 0;
#pragma unroll
        for (int l = 0; l < WL; l++) {
            const unsigned long long v = ((unsigned long long)scratch[NM * WL + tid * WL + l] << 32 | scratch[tid * WL + l]) + carry;
            publish_word(out.result + tid * WL + l, out.seq, (uint32_t)v);
            carry = v >> 32;
        }
    }
    if (tid == 0) *out.ticket = 0;
    GKR_T(6);
    wait_challenge(out.chal);
}

// PAR = 1 layout: per-thread WL-limb accumulators in shared memory (sm[(k*WL+l)*BLOCK + tid]) -> block tree -> grid stage
template <int NM, int WL, int BLOCK>
__device__ __forceinline__ void grid_reduce_wide_raw(uint32_t* sm, const WideOut& out) {
    const int tid = threadIdx.x;
    __syncthreads();
    GKR_T(4);
#pragma unroll 1
    for (int stride = BLOCK / 2; stride >= 1; stride >>= 1) {
        const int items = NM * stride;
#pragma unroll 1
        for (int it = tid; it < items; it += BLOCK) {
            const int k = it / stride, i = it - k * stride;
            widen_pair_add<WL, BLOCK>(sm, k, i, stride);
        }
        __syncthreads();
    }
    // compact the block totals (column 0 of every row) to the front, then reuse the rest as scratch
    uint32_t v[(NM * WL + BLOCK - 1) / BLOCK];
#pragma unroll
    for (int j = 0; j < (NM * WL + BLOCK - 1) / BLOCK; j++) {
        const int i = tid + j * BLOCK;
        v[j] = i < NM * WL ? sm[i * BLOCK] : 0;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < (NM * WL + BLOCK - 1) / BLOCK; j++) {
        const int i = tid + j * BLOCK;
        if (i < NM * WL) sm[i] = v[j];
    }
    __syncthreads();
    grid_stage_wide<NM, WL, BLOCK>(sm, sm + NM * WL, out);
}

// INL layout: 16-limb accumulators in shared memory (sm[(k*16+l)*BLOCK + tid]) + each thread's overflow counts in the bytes of
// `ovf`.  The block tree adds 512-bit numbers, so it too can carry out of limb 15: those carries join the executing thread's
// counts, which are then summed over the block (warp shuffles + shared atomics) into limb 16 of the totals.
template <int NM, int BLOCK>
__device__ __forceinline__ void grid_reduce_wide_raw16(uint32_t* sm, unsigned long long ovf, const WideOut& out) {
    const int tid = threadIdx.x;
    __shared__ unsigned int s_cnt[8];
    if (tid < 8) s_cnt[tid] = 0;
    __syncthreads();
    GKR_T(4);
#pragma unroll 1
    for (int stride = BLOCK / 2; stride >= 1; stride >>= 1) {
        const int items = NM * stride;
#pragma unroll 1
        for (int it = tid; it < items; it += BLOCK) {
            const int k = it / stride, i = it - k * stride;
            uint32_t x[16], y[16];
#pragma unroll
            for (int l = 0; l < 16; l++) {
                x[l] = sm[(k * 16 + l) * BLOCK + i];
                y[l] = sm[(k * 16 + l) * BLOCK + i + stride];
            }
            uint32_t c = add8_carry(x, y);
            c = add8_carry_in(x + 8, y + 8, c);
            ovf += (unsigned long long)c << (8 * k);
#pragma unroll
            for (int l = 0; l < 16; l++) sm[(k * 16 + l) * BLOCK + i] = x[l];
        }
        __syncthreads();
    }
#pragma unroll
    for (int k = 0; k < NM; k++) {
        unsigned int c = (unsigned int)(ovf >> (8 * k)) & 0xffu;
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) c += __shfl_xor_sync(0xffffffffu, c, off);
        if ((tid & 31) == 0 && c) atomicAdd(&s_cnt[k], c);
    }
    // compact the block totals (column 0 of every row + the counts) to the front as 17-limb numbers, then reuse the rest as scratch
    constexpr int PER = (NM * 16 + BLOCK - 1) / BLOCK;
    uint32_t v[PER];
#pragma unroll
    for (int j = 0; j < PER; j++) {
        const int i = tid + j * BLOCK;
        v[j] = i < NM * 16 ? sm[i * BLOCK] : 0;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < PER; j++) {
        const int i = tid + j * BLOCK;
        if (i < NM * 16) sm[(i >> 4) * 17 + (i & 15)] = v[j];
    }
    if (tid < NM) sm[tid * 17 + 16] = s_cnt[tid];
    __syncthreads();
    grid_stage_wide<NM, 17, BLOCK>(sm, sm + NM * 17, out);
}

// Identity gate (layer 2 of the MiMC circuit): the three products eq_t * X_t only feed the round sums, so they are accumulated as
// PLAIN 512-bit products (64 instead of 136 wide multiply-adds each; the host applies one REDC per sum, exactly like the
// factored cipher round) -- at 128 B per pair that moves the kernel from the multiplier's side of B200's ridge to the HBM side.
template <int GATE, bool FOLD, int BLOCK, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB) k_round(const RoundArgs a) {
    constexpr int NEV = GateTraits<GATE>::NEV;
    constexpr bool WIDE = GATE == GATE_IDENTITY;
    constexpr int WL = WIDE ? 17 : 9;
    extern __shared__ uint32_t sm[];  // NEV * WL * BLOCK words
    const int tid = threadIdx.x;
#pragma unroll 1
    for (int i = tid; i < NEV * WL * BLOCK; i += BLOCK) sm[i] = 0;  // own columns only (i % BLOCK == tid)
    const Fr r = a.r_dev ? fr_load(a.r_dev) : fr_unpack(a.r);

    for (size_t x = (size_t)blockIdx.x * BLOCK + tid; x < a.half; x += (size_t)gridDim.x * BLOCK) {
        Fr e, de, s, ds;
        {
            Fr e1;
            load_pair<FOLD>(a.src[0], a.dst[0], x, a.half, r, e, e1);
            de = fr_sub(e1, e);
        }
        if (GATE == GATE_CIPHER) {
            // s_t = X0_t + X1_t + ark  (cipher.go:34-35): sum bottom and top values over the two input tables
            Fr s1 = fr_unpack(a.ark);
            s = s1;
#pragma unroll 1
            for (int tb = 1; tb <= 2; tb++) {
                Fr b, t;
                load_pair<FOLD>(a.src[tb], a.dst[tb], x, a.half, r, b, t);
                s = fr_add(s, b);
                s1 = fr_add(s1, t);
            }
            ds = fr_sub(s1, s);
        } else {
            Fr s1;
            load_pair<FOLD>(a.src[1], a.dst[1], x, a.half, r, s, s1);
            ds = fr_sub(s1, s);
        }
        // evals[t] += eq_t * gate(s_t), t = 0..NEV-1; (e,s) advance by the differences (algo.go:149-199)
#pragma unroll 1
        for (int t = 0; t < NEV; t++) {
            if (WIDE) {
                fr_mul_acc_wide(sm + (size_t)t * WL * BLOCK + tid, BLOCK, e, s);
            } else {
                const Fr term = fr_mul(e, fr_pow7(s));
                wide_acc_add<BLOCK>(sm + (size_t)t * 9 * BLOCK + tid, term);
            }
            e = fr_add(e, de);
            s = fr_add(s, ds);
        }
    }
    if (WIDE) {
        const WideOut wo{a.partials_w, a.red.ticket, a.red.result, a.red.seq, a.red.chal};
        grid_reduce_wide_raw<NEV, 17, BLOCK>(sm, wo);
    } else {
        grid_reduce_wide<NEV, BLOCK>(sm, a.red);
    }
}

// Suffix eq tables of one layer's challenge vector q[0..n): block 0 builds the stages of the low part
// v = q[n-c : n] (seeded with `seed`, the multi-GPU shard factor), block 1 those of the high part v = q[1 : n-c].
// Stage j (j = 0..nv) = eq(last j variables of v, .) with 2^j entries at out[2^j .. 2^(j+1)), built by doubling
// from the LAST variable (same recurrence as poly/eq.go:49-56, applied back to front so that every stage is kept).
struct EqSuffixArgs {
    FrRaw q[28];
    FrRaw seed;
    int n, c;
    FrRaw* outB;  // 2^(c+1) entries
    FrRaw* outA;  // 2^(n-c) entries
};
__global__ void __launch_bounds__(256) k_eq_suffix(const EqSuffixArgs a) {
    const bool isA = blockIdx.x == 1;
    const int nv = isA ? a.n - a.c - 1 : a.c;
    const FrRaw* v = isA ? a.q + 1 : a.q + (a.n - a.c);
    FrRaw* out = isA ? a.outA : a.outB;
    if (threadIdx.x == 0) fr_store(out + 1, isA ? fr_one() : fr_unpack(a.seed));
    __syncthreads();
    for (int j = 0; j < nv; j++) {
        const Fr var = fr_unpack(v[nv - 1 - j]);
        const size_t cnt = (size_t)1 << j;
        for (size_t y = threadIdx.x; y < cnt; y += blockDim.x) {
            const Fr base = fr_load(out + cnt + y);
            const Fr hi = fr_mulc(var, base);
            fr_store(out + 2 * cnt + cnt + y, hi);
            fr_store(out + 2 * cnt + y, fr_sub(base, hi));
        }
        __syncthreads();
    }
}

struct RoundCfArgs {
    const FrRaw* src[2];  // X0, X1 (FOLD: length 4*half, else 2*half)
    FrRaw* dst[2];        // FOLD: folded tables, length 2*half
    size_t half;          // number of pairs x'
    FrRaw r;              // previous challenge (FOLD) ...
    const FrRaw* r_dev;   // ... or, when not null, where the previous launch's last block left it in device memory
    FrConstMul rk;        // KFOLD kernels: the constant-multiplier table of r (80 instead of 136 wide multiply-adds per fold)
    FrRaw ark;            // CipherGate.Ark
    const FrRaw* tA;      // high suffix table of this round or nullptr
    const FrRaw* tB;      // low table (2^c entries when tA != nullptr, else `half` entries)
    int c;
    WideOut red;
};

#ifndef GKR_ACC_LOOP
#define GKR_ACC_LOOP 0
#endif
#ifndef GKR_CUBIC_LOOP
#define GKR_CUBIC_LOOP 0
#endif
// INL: the multiplier is inlined at every call site (big rounds: no call ABI, no IMAD.MOV marshalling on the multiplier's
// pipe) or called out of line (small rounds: short cold instruction fetch)
template <bool INL>
__device__ __forceinline__ Fr cf_mul(const Fr& a, const Fr& b) {
    if (INL) return fr_mul(a, b);
    return fr_mulc(a, b);
}
template <bool INL>
__device__ __forceinline__ Fr cf_sqr(const Fr& a) {
    if (INL) return fr_sqr(a);
    return fr_mulc(a, a);
}
// INL: the accumulator K keeps 16 limbs in shared memory and its overflow count (limb 16: at most 15 for 2^25 pairs per launch)
// in byte K of the packed register `ovf` -- 448 instead of 476 bytes of shared memory per thread, which is what lets 16 instead of 12
// warps be resident per SM (two 256-thread blocks).  !INL: 17 limbs in shared memory, out-of-line product.
template <bool INL, int K>
__device__ __forceinline__ void cf_acc(uint32_t* sm, int tid, int block, unsigned long long& ovf, const Fr& a, const Fr& b) {
    if (INL) {
        uint32_t* acc = sm + (size_t)K * 16 * block + tid;
        uint32_t p[16];
        fr_mul_wide(p, a, b);
        uint32_t w[16];
#pragma unroll
        for (int l = 0; l < 16; l++) w[l] = acc[l * block];
        uint32_t c = add8_carry(w, p);
        c = add8_carry_in(w + 8, p + 8, c);
        ovf += (unsigned long long)c << (8 * K);
#pragma unroll
        for (int l = 0; l < 16; l++) acc[l * block] = w[l];
    } else {
        fr_mul_acc_wide(sm + (size_t)K * 17 * block + tid, block, a, b);
    }
}

template <bool FOLD, int NM, int PAR, int BLOCK, int MINB, bool INL, bool KFOLD = false>
__global__ void __launch_bounds__(BLOCK, MINB) k_round_cf(const RoundCfArgs a) {
    static_assert(NM == 7 || NM == 8, "7 coefficient sums (m_7 from the claim) or all 8");
    static_assert(PAR == 1 || PAR == 8, "one thread or eight lanes per pair");
    extern __shared__ uint32_t sm[];  // PAR == 1: NM * 17 * BLOCK words; PAR == 8: see CF_SMEM_PAR8
    const int tid = threadIdx.x;
    GKR_T(0);
    constexpr int WL1 = 17;          // limbs of a PAR == 1 sum (plain 512-bit products added up)
    constexpr int WLS = INL ? 16 : 17;  // of which in shared memory
    unsigned long long ovf = 0;      // INL: the 17th limbs, one byte per accumulator
    if (PAR == 1) {
#pragma unroll 1
        for (int i = tid; i < NM * WLS * BLOCK; i += BLOCK) sm[i] = 0;  // own columns only (i % BLOCK == tid)
    }
    GKR_T(1);
    const Fr r = (FOLD && a.r_dev) ? fr_load(a.r_dev) : fr_unpack(a.r);
    const Fr ark = fr_unpack(a.ark);
    const size_t half = a.half, m2 = 2 * half;
    const size_t cmask = ((size_t)1 << a.c) - 1;

    if (PAR == 1) {
        for (size_t x = (size_t)blockIdx.x * BLOCK + tid; x < half; x += (size_t)gridDim.x * BLOCK) {
            // the eq factors are requested first as well (small L1/L2-resident tables, but behind the fold's stores they too waited)
            const Fr ub = fr_load(a.tA ? a.tB + (x & cmask) : a.tB + x);
            Fr ua;
            if (a.tA) ua = fr_load(a.tA + (x >> a.c));
            Fr av, bv;
            if (FOLD) {
                // All eight source entries are requested BEFORE the first fold: the loads and stores are volatile asm, so a
                // load placed after a store stays after it, and the fold-store-load sequence exposed one DRAM latency per
                // folded entry (4 long-scoreboard stall sites per pair in profiles/r1_ncu_k_round_cf_fold_opcode_mix.txt).
                // In place is still safe: a thread reads x, x+half, x+2half, x+3half of each table and writes x, x+half.
                const Fr l0 = fr_load_stream(a.src[0] + x), h0 = fr_load_stream(a.src[0] + x + m2);
                const Fr l1 = fr_load_stream(a.src[0] + x + half), h1 = fr_load_stream(a.src[0] + x + half + m2);
                const Fr l2 = fr_load_stream(a.src[1] + x), h2 = fr_load_stream(a.src[1] + x + m2);
                const Fr l3 = fr_load_stream(a.src[1] + x + half), h3 = fr_load_stream(a.src[1] + x + half + m2);
                const Fr b0 = fr_add(l0, KFOLD ? fr_mul_const(a.rk, fr_sub(h0, l0)) : cf_mul<INL>(r, fr_sub(h0, l0)));
                fr_store(a.dst[0] + x, b0);
                const Fr t0 = fr_add(l1, KFOLD ? fr_mul_const(a.rk, fr_sub(h1, l1)) : cf_mul<INL>(r, fr_sub(h1, l1)));
                fr_store(a.dst[0] + x + half, t0);
                const Fr b1 = fr_add(l2, KFOLD ? fr_mul_const(a.rk, fr_sub(h2, l2)) : cf_mul<INL>(r, fr_sub(h2, l2)));
                fr_store(a.dst[1] + x, b1);
                const Fr t1 = fr_add(l3, KFOLD ? fr_mul_const(a.rk, fr_sub(h3, l3)) : cf_mul<INL>(r, fr_sub(h3, l3)));
                fr_store(a.dst[1] + x + half, t1);
                av = fr_add(fr_add(b0, b1), ark);                 // cipher.go:34-35 on the bottom half
                bv = fr_add(fr_sub(t0, b0), fr_sub(t1, b1));      // (top + ark) - (bottom + ark)
            } else {
                const Fr b0 = fr_load_stream(a.src[0] + x), t0 = fr_load_stream(a.src[0] + x + half);
                const Fr b1 = fr_load_stream(a.src[1] + x), t1 = fr_load_stream(a.src[1] + x + half);
                av = fr_add(fr_add(b0, b1), ark);
                bv = fr_add(fr_sub(t0, b0), fr_sub(t1, b1));
            }
            Fr u = a.tA ? cf_mul<INL>(ua, ub) : ub;
            // m_i = T a^(7-i) b^i as (T * degree-4 monomial) * (degree-3 monomial): the four cubic monomials (6 products),
            // T a, T a^4 and T a b^3 (3 products) give all of m_0..m_6 as products that only feed the sums, and those are
            // accumulated UNREDUCED (fr_mul_acc_wide).  9 Montgomery products per pair instead of 11 for powers + chain.
            Fr v30, v21, v12, v03;
#if GKR_CUBIC_LOOP
            if (INL) {
                // (x, y) = (a, b) then (b, a): x^2, x^2 x, x^2 y -- one rolled body instead of two copies (~9 KB of code); the
                // results rotate through (v30, v21) <- (v03, v12) <- (x^3, x^2 y), so after two turns v30 = a^3, v21 = a^2 b, v03 = b^3, v12 = b^2 a
                Fr x = av, y = bv;
#pragma unroll 1
                for (int k = 0; k < 2; k++) {
                    const Fr x2 = fr_sqr(x);
                    v30 = v03;
                    v21 = v12;
                    v03 = fr_mul(x2, x);
                    v12 = fr_mul(x2, y);
                    const Fr t = x;
                    x = y;
                    y = t;
                }
            } else
#endif
            {
                const Fr a2 = cf_sqr<INL>(av), b2 = cf_sqr<INL>(bv);
                v30 = cf_mul<INL>(a2, av);
                v21 = cf_mul<INL>(a2, bv);
                v12 = cf_mul<INL>(av, b2);
                v03 = cf_mul<INL>(b2, bv);
            }
            if (NM == 8) {  // m_7 = (T b * b^3) * b^3
                const Fr tb4 = cf_mul<INL>(cf_mul<INL>(u, bv), v03);
                cf_acc<INL, 7>(sm, tid, BLOCK, ovf, tb4, v03);
            }
            u = cf_mul<INL>(u, av);  // T a
#if GKR_ACC_LOOP
            if (INL) {
                // The seven plain products as ONE rolled loop body (operands picked with selects on the ALU pipe, which has the
                // headroom): 6 copies fewer of a ~140-instruction body, i.e. ~13 KB less for the instruction cache to stream.
                const Fr u1 = cf_mul<INL>(u, v03);  // T a b^3
                u = cf_mul<INL>(u, v30);            // T a^4
#pragma unroll 1
                for (int k = 0; k < 7; k++) {
                    const int ri = k < 4 ? k : k - 3;  // m_0..m_3 = T a^4 * {a^3, a^2 b, a b^2, b^3};  m_4..m_6 = T a b^3 * {a^2 b, a b^2, b^3}
                    Fr lf, rf;
#pragma unroll
                    for (int l = 0; l < 8; l++) {
                        lf.v[l] = k < 4 ? u.v[l] : u1.v[l];
                        rf.v[l] = ri == 0 ? v30.v[l] : (ri == 1 ? v21.v[l] : (ri == 2 ? v12.v[l] : v03.v[l]));
                    }
                    uint32_t* acc = sm + (size_t)k * 16 * BLOCK + tid;
                    uint32_t p[16];
                    fr_mul_wide(p, lf, rf);
                    uint32_t w[16];
#pragma unroll
                    for (int l = 0; l < 16; l++) w[l] = acc[l * BLOCK];
                    uint32_t c = add8_carry(w, p);
                    c = add8_carry_in(w + 8, p + 8, c);
                    ovf += (unsigned long long)c << (8 * k);
#pragma unroll
                    for (int l = 0; l < 16; l++) acc[l * BLOCK] = w[l];
                }
            } else
#endif
            {
                {
                    const Fr u1 = cf_mul<INL>(u, v03);  // T a b^3
                    cf_acc<INL, 4>(sm, tid, BLOCK, ovf, u1, v21);
                    cf_acc<INL, 5>(sm, tid, BLOCK, ovf, u1, v12);
                    cf_acc<INL, 6>(sm, tid, BLOCK, ovf, u1, v03);
                }
                u = cf_mul<INL>(u, v30);  // T a^4
                cf_acc<INL, 3>(sm, tid, BLOCK, ovf, u, v03);
                cf_acc<INL, 2>(sm, tid, BLOCK, ovf, u, v12);
                cf_acc<INL, 1>(sm, tid, BLOCK, ovf, u, v21);
                cf_acc<INL, 0>(sm, tid, BLOCK, ovf, u, v30);
            }
        }
    } else {
        const int j = tid & 7;
        const size_t pstride = (size_t)gridDim.x * (BLOCK / 8);
        uint32_t wacc[9];  // this lane's 288-bit sum of term j
#pragma unroll
        for (int l = 0; l < 9; l++) wacc[l] = 0;
        // block-uniform trip count (the shuffles below need every lane of the warp); surplus lanes redo pair 0 unrecorded
        for (size_t xb = (size_t)blockIdx.x * (BLOCK / 8); xb < half; xb += pstride) {
            const bool live = xb + (tid >> 3) < half;
            const size_t x = live ? xb + (tid >> 3) : 0;
            // lanes j and j+4 both form folded value (j & 3): 0 = X0 bottom, 1 = X0 top, 2 = X1 bottom, 3 = X1 top
            const int tb = (j >> 1) & 1, top = j & 1;
            const size_t idx = x + (top ? half : 0);
            Fr f;
            if (FOLD) {
                const Fr lo = fr_load_stream(a.src[tb] + idx), hi = fr_load_stream(a.src[tb] + idx + m2);
                f = fr_add(lo, fr_mulc(r, fr_sub(hi, lo)));
                __syncwarp();  // in-place fold: all lanes have read their sources before any lane overwrites them
                if (live && j < 4) fr_store(a.dst[tb] + idx, f);
            } else {
                f = fr_load_stream(a.src[tb] + idx);
            }
            Fr u = a.tA ? fr_mulc(fr_load(a.tA + (x >> a.c)), fr_load(a.tB + (x & cmask))) : fr_load(a.tB + x);
            Fr f0, f1, f2, f3;
            const int base = (threadIdx.x & 31) & ~7;
#pragma unroll
            for (int l = 0; l < 8; l++) {
                f0.v[l] = __shfl_sync(0xffffffffu, f.v[l], base + 0);
                f1.v[l] = __shfl_sync(0xffffffffu, f.v[l], base + 1);
                f2.v[l] = __shfl_sync(0xffffffffu, f.v[l], base + 2);
                f3.v[l] = __shfl_sync(0xffffffffu, f.v[l], base + 3);
            }
            GKR_T(2);
            const Fr av = fr_add(fr_add(f0, f2), ark);
            const Fr bv = fr_add(fr_sub(f1, f0), fr_sub(f3, f2));
            // term j = T * a^(7-j) * b^j : bit i of j selects b^(2^i), otherwise a^(2^i); every lane squares its OWN base
            // (lanes with bit i set only ever need powers of b for that bit), so the ladder is 2 squarings + 3 products deep
            Fr s1, s2, s4;
#pragma unroll
            for (int l = 0; l < 8; l++) s1.v[l] = (j & 1) ? bv.v[l] : av.v[l];
            Fr w2;
#pragma unroll
            for (int l = 0; l < 8; l++) w2.v[l] = (j & 2) ? bv.v[l] : av.v[l];
            s2 = fr_sqrc(w2);
            Fr w4;
#pragma unroll
            for (int l = 0; l < 8; l++) w4.v[l] = (j & 4) ? bv.v[l] : av.v[l];
            s4 = fr_sqrc(fr_sqrc(w4));
            u = fr_mulc(fr_mulc(fr_mulc(u, s1), s2), s4);
            GKR_T(3);
            if (live && j < NM) {
                uint32_t t9[9];
#pragma unroll
                for (int l = 0; l < 8; l++) t9[l] = u.v[l];
                t9[8] = 0;
                wide9_add(wacc, t9);
            }
        }
        // lanes j, j+8, j+16, j+24 hold the same accumulator: fold them, then the warps through shared memory
#pragma unroll
        for (int off = 8; off <= 16; off <<= 1) {
            uint32_t o[9];
#pragma unroll
            for (int l = 0; l < 9; l++) o[l] = __shfl_xor_sync(0xffffffffu, wacc[l], off);
            wide9_add(wacc, o);
        }
        constexpr int NW = BLOCK / 32;
        uint32_t* tot = sm;                  // [8][9] block totals
        uint32_t* stage = sm + 8 * 9;        // [NW][8][9]
        const int warp = tid >> 5, lane = tid & 31;
        GKR_T(4);
        if (lane < 8) {
#pragma unroll
            for (int l = 0; l < 9; l++) stage[(warp * 8 + lane) * 9 + l] = wacc[l];
        }
        __syncthreads();
        if (tid < 8) {
            uint32_t acc9[9];
#pragma unroll
            for (int l = 0; l < 9; l++) acc9[l] = stage[tid * 9 + l];
#pragma unroll
            for (int w = 1; w < NW; w++) {
                uint32_t o[9];
#pragma unroll
                for (int l = 0; l < 9; l++) o[l] = stage[(w * 8 + tid) * 9 + l];
                wide9_add(acc9, o);
            }
#pragma unroll
            for (int l = 0; l < 9; l++) tot[tid * 9 + l] = acc9[l];
        }
        __syncthreads();
        grid_stage_wide<NM, 9, BLOCK>(tot, sm + 8 * 9 + NW * 8 * 9, a.red);
        return;
    }
    if (INL) grid_reduce_wide_raw16<NM, BLOCK>(sm, ovf, a.red);
    else grid_reduce_wide_raw<NM, WL1, BLOCK>(sm, a.red);
}

// copies n tagged 64-bit words (see publish_word) to mapped host memory; the tags travel with the data
__global__ void k_publish_words(const unsigned long long* __restrict__ src, int n, unsigned long long* dst) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const unsigned long long v = src[i];
        asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(dst + i), "l"(v) : "memory");
    }
}

// tags n 32-bit words with seq and publishes them (residual tables of one rank -> its slot of the exchange window)
__global__ void k_publish_tagged(const uint32_t* __restrict__ src, int n, unsigned long long* dst, uint32_t seq) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) publish_word(dst + i, seq, src[i]);
}

// ------------------------------------------------------------------------------------------------
// Multi-GPU helpers
// ------------------------------------------------------------------------------------------------
// de-interleave a contiguous slice by owner rank: dst[d*(n/G) + j] = src[j*G + d]  (block d is what rank d receives)
__global__ void __launch_bounds__(256) k_destripe(const FrRaw* __restrict__ src, FrRaw* __restrict__ dst, size_t n, int G) {
    const size_t blk = n / (size_t)G;
    for (size_t o = (size_t)blockIdx.x * blockDim.x + threadIdx.x; o < n; o += (size_t)gridDim.x * blockDim.x) {
        const size_t d = o / blk, j = o - d * blk;
        fr_store(dst + o, fr_load_stream(src + j * (size_t)G + d));
    }
}
// strided shard of a full table: dst[j] = src[j*G + g]
__global__ void __launch_bounds__(256) k_take_shard(const FrRaw* __restrict__ src, FrRaw* __restrict__ dst, size_t n_local, int G, int g) {
    for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < n_local; j += (size_t)gridDim.x * blockDim.x)
        fr_store(dst + j, fr_load_stream(src + j * (size_t)G + g));
}
// ------------------------------------------------------------------------------------------------
// Hint I/O (prover/gadget/hints.go:197-233): batched Montgomery <-> regular conversion and the user-visible hash
// ------------------------------------------------------------------------------------------------
// out[i] = to_mont ? in[i]*2^256 mod q (fr.Element.SetBigInt of a reduced value) : in[i]*2^-256 mod q (ToBigIntRegular)
__global__ void __launch_bounds__(256) k_convert(const FrRaw* __restrict__ in, FrRaw* __restrict__ out, size_t n, int to_mont) {
    Fr f = fr_zero();
    if (to_mont) {  // R^2 mod q
        f.v[0] = 0xae216da7u; f.v[1] = 0x1bb8e645u; f.v[2] = 0xe35c59e3u; f.v[3] = 0x53fe3ab1u;
        f.v[4] = 0x53bb8085u; f.v[5] = 0x8c49833du; f.v[6] = 0x7f4e44a5u; f.v[7] = 0x0216d0b1u;
    } else {
        f.v[0] = 1;
    }
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        fr_store(out + i, fr_mulc(fr_load_stream(in + i), f));
}
// out[x] = a93[x] (+ 2*key[x] + msg[x] when hash != 0: hash.MimcUpdateInplace, hash/mimc.go:24-28 / gadget_api.go:28), optionally
// in regular form
__global__ void __launch_bounds__(256) k_hash_out(const FrRaw* __restrict__ a93, const FrRaw* __restrict__ key, const FrRaw* __restrict__ msg,
                                                   FrRaw* __restrict__ out, size_t n, int hash, int regular) {
    Fr one_raw = fr_zero();
    one_raw.v[0] = 1;
    for (size_t x = (size_t)blockIdx.x * blockDim.x + threadIdx.x; x < n; x += (size_t)gridDim.x * blockDim.x) {
        Fr v = fr_load_stream(a93 + x);
        if (hash) {
            const Fr k = fr_load_stream(key + x);
            v = fr_add(fr_add(v, fr_dbl(k)), fr_load_stream(msg + x));
        }
        if (regular) v = fr_mulc(v, one_raw);
        fr_store(out + x, v);
    }
}

// ------------------------------------------------------------------------------------------------
// element-wise ops for arithmetic parity tests
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_fr_batch(int op, const FrRaw* __restrict__ a, const FrRaw* __restrict__ b, FrRaw* __restrict__ out, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const Fr x = fr_load(a + i);
        Fr y = fr_zero();
        if (op != 3 && op != 4) y = fr_load(b + i);
        Fr z;
        switch (op) {
            case 0: z = fr_mul(x, y); break;
            case 1: z = fr_add(x, y); break;
            case 2: z = fr_sub(x, y); break;
            case 4: z = fr_sqr(x); break;
            default: z = fr_pow7(x); break;
        }
        fr_store(out + i, z);
    }
}

// ------------------------------------------------------------------------------------------------
// Integer-pipe microbenchmarks (roofline denominators)
// ------------------------------------------------------------------------------------------------
// kind 0: 8 IMAD.WIDE.U32 accumulation chains per thread.  Every multiply takes one operand from the NEIGHBOUR chain's
// previous value, so no product is loop-invariant: an earlier version multiplied the same x*y every iteration and ptxas
// strength-reduced the whole loop to IADD3 (it reported 18.5 T "MAC"/s that were additions).  Honest rate on B200:
// ~7-9 T/s (a 32x32->64 multiply-add issues at half the 32-bit IMAD rate).
__global__ void __launch_bounds__(256) k_bench_imad_wide(uint64_t* out, int iters, uint32_t seed) {
    uint64_t c[8];
#pragma unroll
    for (int i = 0; i < 8; i++) c[i] = (uint64_t)seed * (i + 1) + threadIdx.x;
    const uint32_t y = seed | 1u;
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
#pragma unroll
            for (int k = 0; k < 8; k++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(c[k]) : "r"((uint32_t)c[(k + 1) & 7]), "r"(y));
        }
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = c[0] ^ c[1] ^ c[2] ^ c[3] ^ c[4] ^ c[5] ^ c[6] ^ c[7];
}
// kind 1: two independent dependent-chains of fr_mul per thread (what the prover kernels look like)
template <int MULT>  // 0: the library's multiplier (fr_mul), 1: the out-of-line copy (fr_mulc)
__device__ __forceinline__ Fr bench_mul(const Fr& a, const Fr& b) {
    return MULT == 1 ? fr_mulc(a, b) : fr_mul(a, b);
}
template <int MULT>
__global__ void k_bench_fr_mul(FrRaw* out, int iters, uint32_t seed) {
    extern __shared__ uint32_t sm_dummy[];  // only used to limit occupancy from the host side
    Fr a = fr_one(), b = fr_one();
    a.v[0] ^= seed + threadIdx.x;
    b.v[1] ^= seed + blockIdx.x;
    a = fr_reduce_once(a);
    b = fr_reduce_once(b);
    Fr c = b, d = a;
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
        a = bench_mul<MULT>(a, b);
        c = bench_mul<MULT>(c, d);
    }
    fr_store(out + (size_t)blockIdx.x * blockDim.x + threadIdx.x, fr_add(a, c));
}
// kind 3: a single dependent chain per thread
__global__ void k_bench_fr_mul1(FrRaw* out, int iters, uint32_t seed) {
    extern __shared__ uint32_t sm_dummy[];
    Fr a = fr_one(), b = fr_one();
    a.v[0] ^= seed + threadIdx.x;
    b.v[1] ^= seed + blockIdx.x;
    a = fr_reduce_once(a);
    b = fr_reduce_once(b);
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
        a = fr_mul(a, b);
        a = fr_mul(a, b);
    }
    fr_store(out + (size_t)blockIdx.x * blockDim.x + threadIdx.x, a);
}
// kind 2: carry-chained IMAD.WIDE.U32.X: four independent 4-column chains per step (16 wide MACs)
__global__ void __launch_bounds__(256) k_bench_imad_wide_x(uint64_t* out, int iters, uint32_t seed) {
    uint32_t c[4][9];
#pragma unroll
    for (int k = 0; k < 4; k++)
#pragma unroll
        for (int i = 0; i < 9; i++) c[k][i] = seed * (k + 1) + i + threadIdx.x;
    const uint32_t x0 = seed ^ threadIdx.x, x1 = x0 * 3u, x2 = x0 * 5u, x3 = x0 * 7u, y = seed + blockIdx.x;
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
#pragma unroll
            for (int k = 0; k < 4; k++)
                chain4(c[k][0], c[k][1], c[k][2], c[k][3], c[k][4], c[k][5], c[k][6], c[k][7], c[k][8], x0, x1, x2, x3, y);
        }
    }
    uint64_t r = 0;
#pragma unroll
    for (int k = 0; k < 4; k++)
#pragma unroll
        for (int i = 0; i < 9; i++) r += c[k][i];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = r;
}

}  // namespace gkr
