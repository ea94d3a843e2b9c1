// induce_sym.cuh -- symmetric pair-tile Vatistas induction (uniform core radius), sm_100a FP64.
//
// The reference's mutual_ind (src/lowOrder2D/calcs.jl:491-510) evaluates each unordered pair
// once and updates both vortices; with a uniform core radius the two kernel magnitudes are
// equal (magitr == magjtr, :499-500), so one rsqrt feeds two interactions.  This kernel keeps
// that saving on the GPU:
//
//   * vortices are cut into blocks of TB = 256*R; the upper triangle of block pairs (I <= J) is
//     linearised row-major, every pair tile is cut into kSymSub = 32 sub-units (quarters of the 8 phases
//     of the Latin square below / 32 column slices of a diagonal tile), and the list of sub-units is
//     cut into equal contiguous chunks, one per persistent CTA (and per rank first, when the pair work
//     is split over GPUs) -- the load balance is 1/32 tile fine, which matters while there are only
//     a few tiles per CTA (N < 10^5, or several GPUs);
//   * rows (block I) live in registers, R per thread; columns (block J) are staged in shared
//     memory by 1-D TMA bulk copies, double-buffered on mbarriers;
//   * inside a tile warp w works on column chunk (w + phase) mod 8 and lane l on column
//     (k + l) mod 128 of that chunk at step k (a Latin square: no two lanes ever touch the same
//     column accumulator in one step, no shuffles, no atomics);
//   * per pair: 16 FP64-pipe instructions for 2 interactions (+2 per R pairs for the
//     shared-memory accumulator) = 8.25 per interaction instead of 13;
//   * diagonal tiles (I == J) and the bound-vortex sources are one-sided (13 per interaction);
//   * every CTA accumulates into its own partial plane P[c][2][n]; it first zeroes the blocks of that
//     plane it is going to touch and records them (SymTouch: two ranges of block indices), so that
//     k_sym_finish adds, in a fixed order, only the planes that hold something for a given block --
//     bit-reproducible, no floating-point atomics, no whole-plane memset / read.
#pragma once
#include "induce.cuh"

namespace unsflow {

constexpr int kSymRMax = 8;
constexpr int kSymTB = kThreads * kSymRMax;        // slab alignment: largest pair-tile block (2048)
constexpr int kSymSub = 32;                        // work units per pair tile (quarters of the 8 phases of the lane/warp rotation)

struct SymArgs {
    const double *x, *z, *g;     // slabs (absolute indexing); dead entries have g = 0
    const Regions *reg;          // device: live ranges; regions [0, nreg) are free-vortex families
    int nreg;
    int bv_reg;                  // index of the bound-vortex region in *reg (one-sided sources) or -1
    double vc4;
    double *P;                   // [C][2][n] partial planes (pu then pw per CTA), then C SymTouch records
    long long n;                 // plane length
    int rank, nranks;            // split of the tile list over GPUs
};

// blocks (indices into the list of TB-wide blocks of the live regions) whose plane entries CTA c wrote:
// [a0, a1) U [b0, b1); everything else in its plane is stale and must not be read
struct SymTouch {
    long long a0, a1, b0, b1;
};
// doubles to allocate for `ctas` planes of length n plus the touch records
inline long long sym_plane_doubles(int ctas, long long n) { return (long long)ctas * 2 * n + (long long)ctas * (long long)(sizeof(SymTouch) / sizeof(double)); }

// sum of the planes that touched each block + 1/2pi, free stream, convection (or the raw sum)
struct SymFinishArgs {
    const double *P;             // planes + touch records written by k_induce_sym
    long long n;                 // plane length
    int ctas, R;                 // number of planes, rows per thread of the launch that wrote them
    const Regions *reg;          // same regions / nreg as the SymArgs
    int nreg;
    double *vx, *vz;             // outputs (absolute indexing, or packed when `packed`)
    double *x, *z;               // positions to convect (may be null)
    const double *fs;            // device pointer to (u_inf, w_inf) or null
    double fs_u, fs_w;
    double dt;                   // 0 -> no convection
    int accumulate;              // 1: vx += result (mutual_ind semantics)
    int raw;                     // 1: store the plain plane sum only
    int packed;                  // raw only: outputs indexed by the live-target ordinal (regions back to back), vz at vx + packed_stride
    long long packed_stride;
};

// contiguous share [begin, end) of `total` work items owned by `part` of `nparts` (ranks, then CTAs)
#ifdef __CUDACC__
__host__ __device__
#endif
inline void split_range(long long total, long long part, long long nparts, long long *begin, long long *end) {
    *begin = total * part / nparts;
    *end = total * (part + 1) / nparts;
}

size_t sym_smem_bytes(int R);
int sym_rows_per_thread(long long n, int nranks = 1);   // rows per thread R (tile = 256*R) chosen for n free vortices split over nranks GPUs
long long sym_threshold();                               // free vortices from which the pair-tile kernel is used
int sym_ctas(int R);                    // persistent CTAs = resident CTAs per SM x SMs (= number of partial planes)
int sym_ctas_max();
int launch_induce_sym(const SymArgs &a, int R, int order, cudaStream_t st);
int launch_sym_finish(const SymFinishArgs &a, long long ntargets_ub, cudaStream_t st);

#if defined(__CUDACC__) && defined(UNSFLOW_INDUCE_KERNELS)

template <int TB>
__device__ __forceinline__ long long sym_blocks_in(long long b, long long e) {
    return e > b ? (e + TB - 1) / TB - b / TB : 0;
}
template <int TB>
__device__ __forceinline__ long long sym_block_start(const Regions &rg, int nreg, long long k) {
    for (int r = 0; r < nreg; r++) {
        const long long n = sym_blocks_in<TB>(rg.begin[r], rg.end[r]);
        if (k < n) return (rg.begin[r] / TB + k) * TB;
        k -= n;
    }
    return -1;
}
// entries of block k that lie inside their region (the rest of the block is dead padding, g = 0)
template <int TB>
__device__ __forceinline__ int sym_block_live(const Regions &rg, int nreg, long long k) {
    for (int r = 0; r < nreg; r++) {
        const long long n = sym_blocks_in<TB>(rg.begin[r], rg.end[r]);
        if (k < n) {
            const long long left = rg.end[r] - (rg.begin[r] / TB + k) * TB;
            return (int)(left < TB ? left : TB);
        }
        k -= n;
    }
    return 0;
}
// row-major upper triangle incl. diagonal: row I starts at I*nb - I*(I-1)/2
__device__ __forceinline__ long long sym_row_offset(long long I, long long nb) { return I * nb - I * (I - 1) / 2; }

// (I, J) of tile t in the row-major upper-triangle list of nb blocks
__device__ __forceinline__ void sym_tile_coords(long long t, long long nb, long long *Io, long long *Jo) {
    long long I = (long long)(((2.0 * nb + 1.0) - sqrt((2.0 * nb + 1.0) * (2.0 * nb + 1.0) - 8.0 * (double)t)) * 0.5);
    if (I < 0) I = 0;
    if (I > nb - 1) I = nb - 1;
    while (I > 0 && sym_row_offset(I, nb) > t) I--;
    while (I + 1 < nb && sym_row_offset(I + 1, nb) <= t) I++;
    *Io = I;
    *Jo = I + (t - sym_row_offset(I, nb));
}

// Blocks this CTA will write: one row -> that row block and its column range; two rows (the usual row crossing: the
// tail of row I, the head of row I+1) -> [I, J1] and [J, nb); more rows -> every block from the first row on.  (A
// crossing CTA used to zero -- and k_sym_finish to read -- everything from its first row on: 1.6 MB of stores at
// 50 blocks, +30 us on the critical CTA of 2 of 8 ranks at N = 10^5, profiles/r02_stepper_rank_shares.md.)  Zero them in the CTA's own plane (the planes are never memset as a whole) and publish the ranges.
// Kept out of line so that the prologue does not perturb the register allocation of the pair loops.
template <int TB>
__device__ __noinline__ void sym_zero_touched(const Regions &reg, int nreg, long long nb, long long t1, long long I, long long J,
                                              double *Pu, double *Pw, SymTouch *touch) {
    long long I1, J1;
    sym_tile_coords(t1 - 1, nb, &I1, &J1);
    SymTouch tc;
    if (I1 == I) tc = SymTouch{I, I + 1, J, J1 + 1};                 // one row: its row block, columns J..J1
    else if (I1 == I + 1) tc = SymTouch{I, J1 + 1, J, nb};          // two rows: row blocks I, I+1, columns I+1..J1 of the second, J..nb-1 of the first
    else tc = SymTouch{I, nb, 0, 0};                                 // more rows: everything from the first row on
    const int tid = threadIdx.x;
    if (tid == 0) *touch = tc;
    for (int part = 0; part < 2; part++) {
        const long long kb = part ? tc.b0 : tc.a0, ke = part ? tc.b1 : tc.a1;
        for (long long k = kb; k < ke; k++) {
            if (part && k >= tc.a0 && k < tc.a1) continue;
            const long long e0 = sym_block_start<TB>(reg, nreg, k);
            double2 *pu2 = reinterpret_cast<double2 *>(Pu + e0), *pw2 = reinterpret_cast<double2 *>(Pw + e0);
            for (int c = tid; c < TB / 2; c += kThreads) { pu2[c] = make_double2(0.0, 0.0); pw2[c] = make_double2(0.0, 0.0); }
        }
    }
}

// row sums of the finished row block into the CTA's plane: all loads first (one L2 round trip, not R dependent ones)
template <int R>
__device__ __forceinline__ void sym_flush_rows(double *pu_t, double *pw_t, const double (&u)[R], const double (&w)[R]) {
    double pu[R], pw[R];
#pragma unroll
    for (int r = 0; r < R; r++) { pu[r] = pu_t[r * kThreads]; pw[r] = pw_t[r * kThreads]; }
#pragma unroll
    for (int r = 0; r < R; r++) { pu_t[r * kThreads] = pu[r] + u[r]; pw_t[r * kThreads] = pw[r] + w[r]; }
}

template <int R, int ORDER>
__global__ void __launch_bounds__(kThreads, (R >= 8 ? 1 : (R >= 4 ? 2 : 3))) k_induce_sym(const SymArgs a) {
    constexpr int TB = kThreads * R, kChunk = TB / (kThreads / 32);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *s_col = reinterpret_cast<double *>(smem_raw);             // [2][3][TB]
    double2 *s_acc = reinterpret_cast<double2 *>(s_col + 2 * 3 * TB); // [TB]
    uint64_t *s_bar = reinterpret_cast<uint64_t *>(s_acc + TB);       // [2]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const Regions reg = *a.reg;
    long long nb = 0;
    for (int r = 0; r < a.nreg; r++) nb += sym_blocks_in<TB>(reg.begin[r], reg.end[r]);
    const long long W = nb * (nb + 1) / 2;
    // rank range, then this CTA's contiguous chunk of it, in sub-units of 1/kSymSub tile
    long long r0, r1, c0, c1;
    split_range(W * kSymSub, a.rank, a.nranks, &r0, &r1);
    split_range(r1 - r0, blockIdx.x, gridDim.x, &c0, &c1);
    const long long u0 = r0 + c0, u1 = r0 + c1;
    SymTouch *touch = reinterpret_cast<SymTouch *>(a.P + (long long)gridDim.x * 2 * a.n) + blockIdx.x;
    if (u0 >= u1) {
        if (tid == 0) *touch = SymTouch{0, 0, 0, 0};
        return;
    }
    const long long t0 = u0 / kSymSub, t1 = (u1 + kSymSub - 1) / kSymSub;

    long long I, J;
    sym_tile_coords(t0, nb, &I, &J);

    double *Pu = a.P + (long long)blockIdx.x * 2 * a.n;
    double *Pw = Pu + a.n;
    sym_zero_touched<TB>(reg, a.nreg, nb, t1, I, J, Pu, Pw, touch);
    if (tid == 0) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        mbar_fence_init();
    }
    __syncthreads();

    auto issue = [&](long long Jblk, int stage) {
        const long long e0 = sym_block_start<TB>(reg, a.nreg, Jblk);
        constexpr uint32_t bytes = TB * sizeof(double);
        double *dst = s_col + stage * 3 * TB;
        mbar_expect_tx(&s_bar[stage], 3 * bytes);
        tma_load_1d(dst, a.x + e0, bytes, &s_bar[stage]);
        tma_load_1d(dst + TB, a.z + e0, bytes, &s_bar[stage]);
        tma_load_1d(dst + 2 * TB, a.g + e0, bytes, &s_bar[stage]);
    };
    // A block with at most one column chunk of live entries (a short vortex family next to a long one:
    // 17 LEVs beside 10^5 TEVs) is always taken as the COLUMN side of its tiles, where the work can be
    // cut down to the live columns; `rb` / `cb` = row / column block of tile (I, J).
    auto roles = [&](long long Ib, long long Jb, long long *rb, long long *cb, int *ncol) {
        const int li = sym_block_live<TB>(reg, a.nreg, Ib), lj = sym_block_live<TB>(reg, a.nreg, Jb);
        const bool swap = Ib != Jb && li <= kChunk && li < lj;
        *rb = swap ? Jb : Ib;
        *cb = swap ? Ib : Jb;
        *ncol = swap ? li : lj;
    };
    long long rb, cb;
    int ncol;
    roles(I, J, &rb, &cb, &ncol);
    if (tid == 0) issue(cb, 0);

    double tx[R], tz[R], tg[R], u[R], w[R];
    long long rowI = -1, row_start = 0;
    const double vc4 = a.vc4;

    for (long long t = t0; t < t1; t++) {
        const int k = (int)(t - t0), stage = k & 1;
        // next tile in row-major order
        long long In = I, Jn = J + 1;
        if (Jn >= nb) { In = I + 1; Jn = In; }
        long long rbn = 0, cbn = 0;
        int ncoln = 0;
        if (t + 1 < t1) roles(In, Jn, &rbn, &cbn, &ncoln);
        if (tid == 0 && t + 1 < t1) issue(cbn, stage ^ 1);

        if (rb != rowI) {
            if (rowI >= 0) sym_flush_rows<R>(Pu + row_start + tid, Pw + row_start + tid, u, w);
            rowI = rb;
            row_start = sym_block_start<TB>(reg, a.nreg, rb);
#pragma unroll
            for (int r = 0; r < R; r++) {
                const long long i = row_start + tid + r * kThreads;
                tx[r] = a.x[i]; tz[r] = a.z[i]; tg[r] = a.g[i];
                u[r] = 0.0; w[r] = 0.0;
            }
        }
        const double *sx = s_col + stage * 3 * TB, *sz = sx + TB, *sg = sx + 2 * TB;
        // sub-units [pb, pe) of this tile belong to this CTA (all of them except on its first / last tile)
        const int pb = t == t0 ? (int)(u0 - t0 * kSymSub) : 0;
        const int pe = t == t1 - 1 ? (int)(u1 - (t1 - 1) * kSymSub) : kSymSub;
        if (I == J) {
            // ---- diagonal tile: one-sided, broadcast column reads (self pair is an exact 0) ----
            mbar_wait(&s_bar[stage], (k >> 1) & 1);
#pragma unroll 2
            const int jend = min(pe * (TB / kSymSub), ncol);   // columns past the region end are dead
            for (int j = pb * (TB / kSymSub); j < jend; j++) {
                const double xj = sx[j], zj = sz[j], gj = sg[j];
#pragma unroll
                for (int r = 0; r < R; r++) {
                    const double dx = xj - tx[r], dz = zj - tz[r];
                    const double d2 = fma(dz, dz, dx * dx);
                    const double y = rsqrt_full<ORDER>(fma(d2, d2, vc4));
                    const double m = gj * y;
                    u[r] = fma(-dz, m, u[r]);
                    w[r] = fma(dx, m, w[r]);
                }
            }
            if (a.bv_reg >= 0 && pb == 0) {   // bound vortices act on this row block exactly once (here)
                // staged through the column accumulators' shared memory (free on a diagonal tile) in one coalesced pass:
                // read one by one from global memory every bound vortex cost an exposed L2 round trip (69 of them = 35 us,
                // 5 % of a CTA's share when 8 GPUs split a 10^5 wake -- profiles/r02_stepper_rank_shares.md)
                double *sb = reinterpret_cast<double *>(s_acc);
                constexpr int cap = (2 * TB) / 3;
                const long long b0 = reg.begin[a.bv_reg], b1 = reg.end[a.bv_reg];
                for (long long base = b0; base < b1; base += cap) {
                    const int nbv = (int)(b1 - base < cap ? b1 - base : cap);
                    __syncthreads();
                    for (int c = tid; c < nbv; c += kThreads) {
                        sb[c] = a.x[base + c]; sb[cap + c] = a.z[base + c]; sb[2 * cap + c] = a.g[base + c];
                    }
                    __syncthreads();
                    for (int j = 0; j < nbv; j++) {
                        const double xj = sb[j], zj = sb[cap + j], gj = sb[2 * cap + j];
#pragma unroll
                        for (int r = 0; r < R; r++) {
                            const double dx = xj - tx[r], dz = zj - tz[r];
                            const double d2 = fma(dz, dz, dx * dx);
                            const double y = rsqrt_full<ORDER>(fma(d2, d2, vc4));
                            const double m = gj * y;
                            u[r] = fma(-dz, m, u[r]);
                            w[r] = fma(dx, m, w[r]);
                        }
                    }
                }
            }
            __syncthreads();
        } else if (ncol <= kChunk) {
            // ---- off-diagonal tile with a single live column chunk: all warps sweep chunk 0 at once,
            //      each into its own accumulator row s_acc[warp][.], only over the steps that touch a live
            //      column.  The whole tile belongs to the CTA that owns its first work unit. ---------------
            if (pb == 0) {
                for (int c = tid; c < TB; c += kThreads) s_acc[c] = make_double2(0.0, 0.0);
                mbar_wait(&s_bar[stage], (k >> 1) & 1);
                __syncthreads();
                double2 *acc_w = s_acc + warp * kChunk;
                for (int s = 0; s < kChunk; s++) {
                    if (s >= ncol && s < kChunk - 31) { s = kChunk - 32; continue; }   // lanes cover columns s..s+31 (mod chunk)
                    const int j = (s + lane) & (kChunk - 1);
                    const double xj = sx[j], zj = sz[j], gj = sg[j];
                    double cu = 0.0, cw = 0.0;
#pragma unroll
                    for (int r = 0; r < R; r++) {
                        const double dx = xj - tx[r], dz = zj - tz[r];
                        const double d2 = fma(dz, dz, dx * dx);
                        const double y = rsqrt_full<ORDER>(fma(d2, d2, vc4));
                        const double t1v = y * dz, t2v = y * dx;
                        u[r] = fma(-gj, t1v, u[r]);
                        w[r] = fma(gj, t2v, w[r]);
                        cu = fma(tg[r], t1v, cu);
                        cw = fma(-tg[r], t2v, cw);
                    }
                    double2 acc = acc_w[j];
                    acc.x += cu;
                    acc.y += cw;
                    acc_w[j] = acc;
                    __syncwarp();
                }
                __syncthreads();
                const long long cstart = sym_block_start<TB>(reg, a.nreg, cb);
                for (int c = tid; c < kChunk; c += kThreads) {
                    double su = 0.0, sw = 0.0;
#pragma unroll
                    for (int q = 0; q < kThreads / 32; q++) { su += s_acc[q * kChunk + c].x; sw += s_acc[q * kChunk + c].y; }
                    Pu[cstart + c] += su;
                    Pw[cstart + c] += sw;
                }
            } else {
                mbar_wait(&s_bar[stage], (k >> 1) & 1);   // keep the barrier phases in step
            }
            __syncthreads();
        } else {
            // ---- off-diagonal tile: every kernel evaluation feeds row i and column j ------------
            for (int c = tid; c < TB; c += kThreads) s_acc[c] = make_double2(0.0, 0.0);
            mbar_wait(&s_bar[stage], (k >> 1) & 1);
            __syncthreads();
            // one phase: warp w works on column chunk (w + p) mod 8, steps [sb, se) of the lane rotation
            auto run_phase = [&](const int p, const int sb, const int se) {
                const int base = ((warp + p) & (kThreads / 32 - 1)) * kChunk;
                // software pipeline: the column read for step s+1 is issued before the
                // accumulator read-modify-write of step s (column data is read-only in a tile)
                int j = base + ((sb + lane) & (kChunk - 1));
                double xj = sx[j], zj = sz[j], gj = sg[j];
#pragma unroll 1
                for (int s = sb; s < se; s++) {
                    const int jn = base + ((s + 1 + lane) & (kChunk - 1));
                    const double xn = sx[jn], zn = sz[jn], gn = sg[jn];
                    double cu = 0.0, cw = 0.0;
#pragma unroll
                    for (int r = 0; r < R; r++) {
                        const double dx = xj - tx[r], dz = zj - tz[r];
                        const double d2 = fma(dz, dz, dx * dx);
                        const double y = rsqrt_full<ORDER>(fma(d2, d2, vc4));
                        const double t1v = y * dz, t2v = y * dx;
                        u[r] = fma(-gj, t1v, u[r]);
                        w[r] = fma(gj, t2v, w[r]);
                        cu = fma(tg[r], t1v, cu);
                        cw = fma(-tg[r], t2v, cw);
                    }
                    double2 acc = s_acc[j];
                    acc.x += cu;
                    acc.y += cw;
                    s_acc[j] = acc;
                    __syncwarp();   // lane l+1 wrote column j+1 in this step; lane l reads it next step
                    j = jn; xj = xn; zj = zn; gj = gn;
                }
                __syncthreads();
            };
            if (pb == 0 && pe == kSymSub) {
                // whole tile (the common case): compile-time loop bounds
#pragma unroll 1
                for (int p = 0; p < kThreads / 32; p++) run_phase(p, 0, kChunk);
            } else {
                // first / last tile of this CTA's share: a unit = kUnit consecutive steps of the
                // (phase, step) sequence of the tile
                constexpr int kUnit = (kThreads / 32) * kChunk / kSymSub;
                const int gs0 = pb * kUnit, gs1 = pe * kUnit;
                for (int p = gs0 / kChunk; p * kChunk < gs1; p++)
                    run_phase(p, max(gs0 - p * kChunk, 0), min(gs1 - p * kChunk, kChunk));
            }
            // flush the column accumulators into this CTA's private plane
            {
                // all plane loads first, then the adds and stores: one L2 round trip instead of R dependent ones
                const long long cstart = sym_block_start<TB>(reg, a.nreg, cb);
                double pu[R], pw[R];
#pragma unroll
                for (int r = 0; r < R; r++) { pu[r] = Pu[cstart + tid + r * kThreads]; pw[r] = Pw[cstart + tid + r * kThreads]; }
#pragma unroll
                for (int r = 0; r < R; r++) {
                    const double2 acc = s_acc[tid + r * kThreads];
                    Pu[cstart + tid + r * kThreads] = pu[r] + acc.x;
                    Pw[cstart + tid + r * kThreads] = pw[r] + acc.y;
                }
            }
            __syncthreads();
        }
        I = In;
        J = Jn;
        rb = rbn; cb = cbn; ncol = ncoln;
    }
    if (rowI >= 0) sym_flush_rows<R>(Pu + row_start + tid, Pw + row_start + tid, u, w);
}

// One CTA = kThreads consecutive entries of one block.  The planes that touched the block are found from the
// SymTouch records (ballot per warp -> ordered list in shared memory) and added in ascending plane order.
template <int R>
__global__ void __launch_bounds__(kThreads) k_sym_finish(const SymFinishArgs a) {
    constexpr int TB = kThreads * R;
    __shared__ unsigned s_mask[16];      // up to 512 planes
    __shared__ short s_list[512];
    __shared__ int s_cnt;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const Regions reg = *a.reg;
    const long long X = blockIdx.x / R;
    long long nb = 0;
    for (int r = 0; r < a.nreg; r++) nb += sym_blocks_in<TB>(reg.begin[r], reg.end[r]);
    if (X >= nb) return;
    const SymTouch *touch = reinterpret_cast<const SymTouch *>(a.P + (long long)a.ctas * 2 * a.n);
    for (int c0 = warp * 32; c0 < a.ctas; c0 += kThreads) {
        const int c = c0 + lane;
        bool in = false;
        if (c < a.ctas) {
            const SymTouch t = touch[c];
            in = (X >= t.a0 && X < t.a1) || (X >= t.b0 && X < t.b1);
        }
        const unsigned m = __ballot_sync(0xffffffffu, in);
        if (lane == 0) s_mask[c0 >> 5] = m;
    }
    __syncthreads();
    if (tid == 0) {
        int cnt = 0;
        for (int wd = 0; wd * 32 < a.ctas; wd++) {
            unsigned m = s_mask[wd];
            while (m) { const int b = __ffs(m) - 1; m &= m - 1; s_list[cnt++] = (short)(wd * 32 + b); }
        }
        s_cnt = cnt;
    }
    __syncthreads();
    // which region does block X belong to, and which entry is mine?
    long long k = X, i = -1, ord = 0;
    for (int r = 0; r < a.nreg; r++) {
        const long long nbr = sym_blocks_in<TB>(reg.begin[r], reg.end[r]);
        if (k < nbr) {
            i = (reg.begin[r] / TB + k) * TB + (long long)(blockIdx.x % R) * kThreads + tid;
            if (i < reg.begin[r] || i >= reg.end[r]) return;
            ord += i - reg.begin[r];
            break;
        }
        k -= nbr;
        ord += reg.end[r] > reg.begin[r] ? reg.end[r] - reg.begin[r] : 0;
    }
    if (i < 0) return;
    const int cnt = s_cnt;
    double su = 0.0, sw = 0.0;
    for (int s = 0; s < cnt; s += 16) {   // masked batches of 16 independent loads; the additions stay in plane order
        double bu[16], bw[16];
#pragma unroll
        for (int q = 0; q < 16; q++) {
            const bool ok = s + q < cnt;
            const long long off = ok ? (long long)s_list[s + q] * 2 * a.n + i : 0;
            bu[q] = ok ? a.P[off] : 0.0;
            bw[q] = ok ? a.P[off + a.n] : 0.0;
        }
#pragma unroll
        for (int q = 0; q < 16; q++) { su += bu[q]; sw += bw[q]; }
    }
    if (a.raw) {
        if (a.packed) { a.vx[ord] = su; a.vx[a.packed_stride + ord] = sw; }
        else { a.vx[i] = su; a.vz[i] = sw; }
        return;
    }
    double vx = su * kInvTwoPi, vz = sw * kInvTwoPi;
    const double fu = a.fs ? a.fs[0] : a.fs_u, fw = a.fs ? a.fs[1] : a.fs_w;
    vx += fu;
    vz += fw;
    if (a.accumulate) { vx += a.vx[i]; vz += a.vz[i]; }
    a.vx[i] = vx;
    a.vz[i] = vz;
    if (a.dt != 0.0 && a.x) {
        a.x[i] += a.dt * vx;
        a.z[i] += a.dt * vz;
    }
}

#endif  // kernels
}  // namespace unsflow
