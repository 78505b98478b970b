// wafer_b200.cu — context, stream orchestration and the C ABI of include/wafer_b200.h.
//
// One wafer_ctx = one GPU = one x-slab of the lattice.  The reference functions each entry point replaces are
// cited in the header; the driver loop `wafer_solve` restates src/grid.rs:50-246 on top of them.
#include "../../include/wafer_b200.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <new>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "generators.cuh"
#include "kernels.cuh"
#include "nccl_dyn.h"
#include "sweep_tb.cuh"
#include "sweep_tma1.cuh"

using namespace wafer;

namespace {
std::string g_create_error;

// scalar slots in ctx->scal (device doubles)
// SL_RAW: [0] = sum psi^2 of the last excited-state sweep, [1 + i] = raw overlap sum q_i psi; SL_COEF: Gram-Schmidt s_i
// SL_CHK: [sum psi^2, sum psi^2 potsub, sum psi^2 r2] left by the last sweep of a ground-state evolve (MODE_CHK)
enum { SL_OBS = 0, SL_TMP = 5, SL_CHK = 8, SL_RAW = 16, SL_COEF = 16 + 256, SL_COUNT = 16 + 512 };
constexpr int GRAM_LD = 256;  // at most 255 stored states (wavenum is a u8)
}  // namespace

struct Tb2Sched {  // work lists of the persistent time-tiled sweep for one plane range (device arrays, sweep_tb.cuh Sched)
    tb::Segment *bulk = nullptr, *segs = nullptr;
    int *first = nullptr, *counters = nullptr;
    int ncta = 0, nbulk = 0;
};

struct wafer_ctx {
    wafer_params p{};
    Geom g{};
    int dev = 0, sm_count = 0;
    int rank = 0, world = 1;
    bool onfly = true;
    bool use_tb = false;            // time-tiled TMA sweep available (ThreePoint, V on the fly)
    CUtensorMap tm_psi[2], tm_v;    // TMA descriptors of the interior of psi[0], psi[1], v
    double* hfield = nullptr;       // h = (dt*v)/2 for the time-tiled sweep (WAFER_TB_HFIELD), rebuilt when V changes
    CUtensorMap tm_h;
    bool h_valid = false;
    int den_ok = 0;
    std::map<std::pair<int, int>, Tb2Sched> tb2_sched;  // keyed by the plane range [xb, xe)
    bool use_t1 = false;            // TMA-pipelined one-step sweep (WAFER_FLAG_TMA_ONE_STEP)
    CUtensorMap t1_psi[2], t1_v, t1_h;
    cudaStream_t s_main = nullptr, s_halo = nullptr;
    cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr, ev_main = nullptr, ev_halo = nullptr;
    double* psi[2] = {nullptr, nullptr};
    int cur = 0;
    double *v = nullptr, *a = nullptr, *b = nullptr, *potsub_arr = nullptr;
    int potsub_mode = 0;
    double potsub_scalar = 0.0;
    std::vector<double*> lowers;
    // CUDA graph of TWO excited-state steps (sweep + coefficients + projection, twice: the ping-pong parity returns), per
    // (wnum, cur): small lattices are launch bound there (three dependent launches of a few microseconds per step)
    std::map<std::pair<int, int>, std::pair<cudaGraphExec_t, int>> step_graphs;  // exec, kernel launches per replay
    bool chk_valid = false;    // scal[SL_CHK..] holds the point-wise check sums of the CURRENT psi (this rank's share)
    double* gram = nullptr;    // G[i][j] = <q_i, q_j>, j < i, of the stored states (device, GRAM_LD x GRAM_LD)
    // host <-> device: two small bounce buffers in the host layout, filled / drained on s_copy while the
    // pack / unpack kernel of the other one runs on s_main (lazily allocated, never a field-sized buffer)
    double* stg[2] = {nullptr, nullptr};
    long long stg_planes = 0;       // padded host planes one bounce buffer holds
    cudaStream_t s_copy = nullptr;
    cudaEvent_t ev_stg_free[2] = {nullptr, nullptr}, ev_stg_full[2] = {nullptr, nullptr};
    double* partials = nullptr;
    long long partials_cap = 0;
    double* scal = nullptr;
    double* h_scal = nullptr;  // pinned mirror for D2H of scalars
    int* ring_flag = nullptr;
    bool have_v = false, have_phi = false;
    uint64_t launches = 0;
    NcclComm comm = nullptr;
    // fused halo over peer memory (CUDA IPC): neighbour 0 = rank-1 (lower x), 1 = rank+1
    bool p2p = false;
    double* peer_psi[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};  // [neighbour][buffer index]
    unsigned long long* flags = nullptr;                  // [0] written by rank-1, [1] by rank+1: passes completed
    unsigned long long* peer_flags[2] = {nullptr, nullptr};
    int* p2p_timeout = nullptr;                           // device flag: a wait gave up (peer died)
    unsigned long long pass = 0;                          // boundary passes completed (same on every rank)
    int peer_L[2] = {0, 0};
    int min_L = 0;                                        // smallest slab over all ranks: uniform overlap decision
    unsigned long long dbg_halo_delay_ns = 0;             // fault injection: stall this rank's halo stream every pass
    unsigned long long* d_cksum = nullptr;
    mutable std::string err;
    size_t bytes() const { return (size_t)g.total() * sizeof(double); }
};

static void drop_step_graphs(wafer_ctx* ctx) {
    for (auto& kv : ctx->step_graphs) cudaGraphExecDestroy(kv.second.first);
    ctx->step_graphs.clear();
}

#define CK(call)                                                                                        \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess) {                                                                        \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_);                             \
            return WAFER_ERR_CUDA;                                                                      \
        }                                                                                               \
    } while (0)

#define NK(call)                                                                                        \
    do {                                                                                                \
        int r_ = (call);                                                                                \
        if (r_ != kNcclSuccess) {                                                                       \
            ctx->err = std::string(#call) + ": " + nccl_api().GetErrorString(r_);                      \
            return WAFER_ERR_NCCL;                                                                      \
        }                                                                                               \
    } while (0)

#define REQUIRE(cond, msg)                                                                              \
    do {                                                                                                \
        if (!(cond)) {                                                                                  \
            ctx->err = msg;                                                                             \
            return WAFER_ERR_INVALID;                                                                   \
        }                                                                                               \
    } while (0)

#define TRY(expr)                                                                                       \
    do {                                                                                                \
        int t_ = (expr);                                                                                \
        if (t_ != WAFER_OK) return t_;                                                                  \
    } while (0)

namespace {

inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

double denominator(const wafer_ctx* c) {
    // grid.rs:569/594/626: r64(k)*dn*dn*mass, left to right
    const double k = c->p.ext == 1 ? 2. : (c->p.ext == 2 ? 24. : 360.);
    return k * c->p.dn * c->p.dn * c->p.mass;
}

int post_launch(wafer_ctx* ctx, int n = 1) {
    ctx->launches += n;
    CK(cudaGetLastError());
    return WAFER_OK;
}

int ew_grid(const wafer_ctx* ctx, long long n2) {
    return (int)std::min<long long>(ceil_div(n2, EW_THREADS), (long long)ctx->sm_count * 8);
}

// reduce `ns` rows of `nblocks` partials into scal[slot..slot+ns) and all-reduce across ranks
int finalize(wafer_ctx* ctx, int ns, int nblocks, int slot, cudaStream_t st) {
    if (ns == 1) finalize_kernel<1><<<1, 1024, 0, st>>>(ctx->partials, nblocks, ctx->scal + slot);
    else finalize_kernel<4><<<1, 1024, 0, st>>>(ctx->partials, nblocks, ctx->scal + slot);
    TRY(post_launch(ctx));
    if (ctx->world > 1)
        NK(nccl_api().AllReduce(ctx->scal + slot, ctx->scal + slot, ns, kNcclFloat64, kNcclSum, ctx->comm, st));
    return WAFER_OK;
}

__global__ void set_scalar_kernel(double* p, double v) { *p = v; }

// ---- sweep dispatch ------------------------------------------------------------------------------------
template <int E>
int launch_sweep_simple(wafer_ctx* ctx, const double* cur, double* nxt, int xb, int xe, bool norm, long long part_off,
                        cudaStream_t st) {
    const Geom& g = ctx->g;
    if (xe <= xb) return WAFER_OK;
    dim3 grid(ceil_div(g.nz, 2 * SW_BX), ceil_div(g.ny, SW_BY), ceil_div(xe - xb, SW_XCH)), block(SW_BX, SW_BY);
    const double den = denominator(ctx), dt = ctx->p.dt;
    double* part = ctx->partials + part_off;
    const double* fa = ctx->onfly ? ctx->v : ctx->a;
    const double* fb = ctx->b;
    if (ctx->onfly) {
        if (norm) sweep_simple_kernel<E, true, true><<<grid, block, 0, st>>>(cur, nxt, fa, fb, g, xb, xe, dt, den, part);
        else sweep_simple_kernel<E, true, false><<<grid, block, 0, st>>>(cur, nxt, fa, fb, g, xb, xe, dt, den, part);
    } else {
        if (norm) sweep_simple_kernel<E, false, true><<<grid, block, 0, st>>>(cur, nxt, fa, fb, g, xb, xe, dt, den, part);
        else sweep_simple_kernel<E, false, false><<<grid, block, 0, st>>>(cur, nxt, fa, fb, g, xb, xe, dt, den, part);
    }
    return post_launch(ctx);
}

int simple_blocks(const wafer_ctx* ctx, int xb, int xe) {
    const Geom& g = ctx->g;
    return ceil_div(g.nz, 2 * SW_BX) * ceil_div(g.ny, SW_BY) * ceil_div(xe - xb, SW_XCH);
}

// forward declarations of the TMA one-step path (defined below, next to the tensor-map helpers)
template <int E> int launch_sweep_t1(wafer_ctx* ctx, int src, int xb, int xe, int mode, cudaStream_t st, int nb_total, int bid_off);
template <int E> int t1_blocks(const wafer_ctx* ctx, int xb, int xe);

// number of per-CTA partial sums the fused-norm sweep over [xb, xe) leaves in ctx->partials
int sweep_blocks(const wafer_ctx* ctx, int xb, int xe);

int launch_sweep(wafer_ctx* ctx, const double* cur, double* nxt, int xb, int xe, int nred, long long part_off,
                 cudaStream_t st, int nb_total = 0, int bid_off = 0);

int launch_sweep_plain(wafer_ctx* ctx, const double* cur, double* nxt, int xb, int xe, bool norm, long long part_off,
                       cudaStream_t st) {
    switch (ctx->p.ext) {
        case 1: return launch_sweep_simple<1>(ctx, cur, nxt, xb, xe, norm, part_off, st);
        case 2: return launch_sweep_simple<2>(ctx, cur, nxt, xb, xe, norm, part_off, st);
        default: return launch_sweep_simple<3>(ctx, cur, nxt, xb, xe, norm, part_off, st);
    }
}

// ---- time-tiled TMA sweep (sweep_tb.cuh): two steps per launch ----------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_tensor_map(wafer_ctx* ctx, CUtensorMap* tm, double* field, int box_rows) {  // 64-column boxes
    static EncodeTiledFn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (!fn || qres != cudaDriverEntryPointSuccess) { ctx->err = "cuTensorMapEncodeTiled not available"; return WAFER_ERR_CUDA; }
        encode = (EncodeTiledFn)fn;
    }
    const Geom& g = ctx->g;
    // the tensor is the lattice interior only: (nz, ny, L+2gx) with the device pitches; everything outside it
    // (y ghost rows, z pads, planes beyond the ghosts) is produced by the TMA zero fill
    cuuint64_t dims[3] = {(cuuint64_t)g.nz, (cuuint64_t)g.ny, (cuuint64_t)(g.L + 2 * g.gx)};
    cuuint64_t strides[2] = {(cuuint64_t)g.zp * sizeof(double), (cuuint64_t)g.plane * sizeof(double)};
    cuuint32_t box[3] = {(cuuint32_t)tb::BW, (cuuint32_t)box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, field + (long long)g.e * g.zp, dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { ctx->err = "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")"; return WAFER_ERR_CUDA; }
    return WAFER_OK;
}

int init_tb(wafer_ctx* ctx) {
    ctx->use_tb = false;
    if (ctx->p.ext != 1 || !ctx->onfly || (ctx->p.flags & WAFER_FLAG_SIMPLE_SWEEP)) return WAFER_OK;
    TRY(make_tensor_map(ctx, &ctx->tm_psi[0], ctx->psi[0], tb::R0));
    TRY(make_tensor_map(ctx, &ctx->tm_psi[1], ctx->psi[1], tb::R0));
    TRY(make_tensor_map(ctx, &ctx->tm_v, ctx->v, tb::R1));
    CK(cudaFuncSetAttribute(tb::sweep_tb2_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tb::SMEM_BYTES));
    CK(cudaFuncSetAttribute(tb::sweep_tb2_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tb::SMEM_BYTES));
    CK(cudaFuncSetAttribute(tb::sweep_tb2_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tb::SMEM_BYTES));
    CK(cudaFuncSetAttribute(tb::sweep_tb2_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tb::SMEM_BYTES));
    const double den = denominator(ctx);
    ctx->den_ok = (den > 7.888609052210118e-31 && den < 1.2676506002282294e30) ? 1 : 0;  // 2^-100 .. 2^100
    ctx->use_tb = true;
    return WAFER_OK;
}

// ---- work distribution of the time-tiled sweep (persistent grid; see sweep_tb.cuh Sched) ---------------------------
// Every segment pays 4 plane-iterations of pipeline fill; CTAs that share a tile edge re-read each other's halo cells,
// from L2 only if they pass the same planes at about the same time; and the SMs should finish together.
//   * x is cut into chunks of at most 512 planes; per chunk, whole tile columns in tile order form the BULK list that the
//     CTAs drain through an atomic counter (in-order dispatch keeps neighbouring tiles close in time);
//   * of the last chunk, the ntiles % G columns that would leave most SMs idle in a last partial wave are cut along x
//     into G equal shares (even lengths, no stubs shorter than 6 planes): the static TAIL of each CTA.
// 512^3 on 148 SMs: one round of 148 columns + 14 columns cut 148 ways = 516 + ~53 plane-iterations per SM, against
// 10 waves x 60 for round 1's uniform grid of 162 tiles x 9 chunks.  Measured (gpurun_out/r2f..r2j,
// profiles/r2_tb2_variants.md): statically assigned rounds +4.8 % at 512^3 and +6 % on a 128-plane slab but -6 % at
// 1024^3, where CTAs drift apart from round to round and the halo re-reads miss L2 (+17 % DRAM reads, ncu r2g); a fully
// contiguous static split ("contig") -13 %.  WAFER_TB_SCHED=dynamic|rounds|legacy|contig selects the policy for A/B runs.
// pure host arithmetic (also behind wafer_tb2_plan for the CPU tests): bulk = in-order list, per[k] = static tail of CTA k
struct Tb2Plan {
    std::vector<tb::Segment> bulk;
    std::vector<std::vector<tb::Segment>> per;
};
Tb2Plan tb2_plan(int ny, int nz, int xb, int xe, int slots, const std::string& mode, int MAXSEG) {
    Tb2Plan pl;
    std::vector<tb::Segment>& bulk = pl.bulk;
    std::vector<std::vector<tb::Segment>>& per = pl.per;
    const int ntz = ceil_div(nz, tb::TZ), nty = ceil_div(ny, tb::TY), ntiles = ntz * nty, P = xe - xb;
    auto seg = [&](int tile, int a, int b) { return tb::Segment{(tile / ntz) * tb::TY, (tile % ntz) * tb::TZ, a, b}; };
    const int MINSEG = 6;
    int G = (int)std::max<long long>(1, std::min<long long>(slots, ((long long)ntiles * P + 7) / 8));
    // shares of `cols` tile columns starting at tile `t0`, planes [ca, cz), cut into G equal runs -> per[k]
    auto cut_static = [&](int t0, int cols, int ca, int cz) {
        const int Pc = cz - ca;
        const long long Ur = (long long)cols * Pc;
        auto snap = [&](long long u) {
            const long long p = u % Pc;
            if (p == 0) return u;
            if (p < MINSEG) return u - p;
            if (Pc - p < MINSEG) return u + (Pc - p);
            return u - (p & 1);
        };
        for (int k = 0; k < G && Ur > 0; ++k) {
            long long u = snap(k * Ur / G);
            const long long ue = k + 1 == G ? Ur : snap((k + 1) * Ur / G);
            while (u < ue) {
                const int p0 = (int)(u % Pc), len = (int)std::min<long long>(Pc - p0, ue - u);
                per[k].push_back(seg(t0 + (int)(u / Pc), ca + p0, ca + p0 + len));
                u += len;
            }
        }
    };
    if (mode == "legacy") {
        // round 1's grid: tiles x chunks CTAs, one segment each, dispatched in order (here through the counter)
        int best_nc = 1;
        double best_cost = 1e300;
        for (int nc = 1; nc <= std::min(P, 64); ++nc) {
            const double cost = (double)ceil_div((long long)ntiles * nc, slots) * ((double)ceil_div(P, nc) + 3.0);
            if (cost < best_cost * 0.999) { best_cost = cost; best_nc = nc; }
        }
        const int chunk = ceil_div(P, best_nc);
        for (int c = 0; c * chunk < P; ++c)
            for (int t = 0; t < ntiles; ++t) bulk.push_back(seg(t, xb + c * chunk, std::min(xe, xb + (c + 1) * chunk)));
        G = (int)std::min<size_t>(slots, bulk.size());
        per.resize(G);
    } else if (mode == "contig" || (mode == "dynamic" && ntiles * 4 < G * 3)) {
        // few tile columns on many CTAs (small planes: every halo lives in L2 anyway): one contiguous static run each
        per.resize(G);
        cut_static(0, ntiles, xb, xe);
    } else {
        if (mode == "dynamic" && ntiles < G) G = ntiles;  // nearly one column per CTA: leave the few spare SMs idle, keep lock step
        per.resize(G);
        // number of x chunks: at least P / MAXSEG; a few more if that leaves a smaller statically cut tail.  The tail runs
        // out of lock step (its halo re-reads miss L2), which the estimate charges with an empirical factor of 1.6.
        int nchunk = ceil_div(P, MAXSEG);
        if (mode == "dynamic" && ntiles >= G) {
            double best = 1e300;
            for (int nc = ceil_div(P, MAXSEG); nc <= ceil_div(P, MAXSEG) + 6 && nc * 32 <= P; ++nc) {
                const int ch = (ceil_div(P, nc) + 1) & ~1;
                const long long items = (long long)nc * ntiles, rest = items % G;
                const double cost = (double)(items / G) * (ch + 4) + (rest ? ((double)rest * ch / G + 4) * 1.6 : 0.0);
                if (cost < best * 0.999) { best = cost; nchunk = nc; }
            }
        }
        const int chunk = (ceil_div(P, nchunk) + 1) & ~1;
        const int R = ntiles / G;
        for (int c = 0; c * chunk < P; ++c) {
            const int ca = xb + c * chunk, cz = std::min(xe, ca + chunk);
            if (mode == "rounds") {  // static assignment of whole columns, round by round
                for (int r = 0; r < R; ++r)
                    for (int k = 0; k < G; ++k) per[k].push_back(seg(r * G + k, ca, cz));
                cut_static(R * G, ntiles - R * G, ca, cz);
            } else {                 // "dynamic": in-order dispatch of whole columns, chunk after chunk
                for (int t = 0; t < ntiles; ++t) bulk.push_back(seg(t, ca, cz));
            }
        }
        if (mode != "rounds") {
            // the bulk must come out even — a multiple of G columns — or the CTAs that drew one column more finish a whole
            // column late: the last bulk.size() % G columns (all in the last chunk) are cut into G equal static shares
            const int rest = std::min((int)(bulk.size() % G), ntiles);  // (never reaches back into an earlier chunk)
            if (rest > 0) {
                const tb::Segment f = bulk[bulk.size() - rest];
                const int t0 = (f.y0 / tb::TY) * ntz + f.z0 / tb::TZ;
                bulk.resize(bulk.size() - rest);
                cut_static(t0, rest, f.xa, f.xz);
            }
        }
    }
    return pl;
}

int tb2_schedule(wafer_ctx* ctx, int xb, int xe, const Tb2Sched** out) {
    auto it = ctx->tb2_sched.find({xb, xe});
    if (it != ctx->tb2_sched.end()) { *out = &it->second; return WAFER_OK; }
    static const std::string mode = getenv("WAFER_TB_SCHED") ? getenv("WAFER_TB_SCHED") : "dynamic";
    static const int MAXSEG = getenv("WAFER_TB_MAXSEG") ? std::max(8, atoi(getenv("WAFER_TB_MAXSEG"))) : 512;
    const Tb2Plan pl = tb2_plan(ctx->g.ny, ctx->g.nz, xb, xe, ctx->sm_count * tb::CTAS_PER_SM, mode, MAXSEG);
    const std::vector<tb::Segment>& bulk = pl.bulk;
    std::vector<tb::Segment> flat;
    std::vector<int> first{0};
    for (const auto& v : pl.per) {
        flat.insert(flat.end(), v.begin(), v.end());
        first.push_back((int)flat.size());
    }
    Tb2Sched sc;
    sc.ncta = (int)pl.per.size();
    sc.nbulk = (int)bulk.size();
    CK(cudaMalloc(&sc.bulk, std::max<size_t>(bulk.size(), 1) * sizeof(tb::Segment)));
    CK(cudaMalloc(&sc.segs, std::max<size_t>(flat.size(), 1) * sizeof(tb::Segment)));
    CK(cudaMalloc(&sc.first, first.size() * sizeof(int)));
    CK(cudaMalloc(&sc.counters, 2 * sizeof(int)));
    CK(cudaMemsetAsync(sc.counters, 0, 2 * sizeof(int), ctx->s_main));
    if (!bulk.empty()) CK(cudaMemcpyAsync(sc.bulk, bulk.data(), bulk.size() * sizeof(tb::Segment), cudaMemcpyHostToDevice, ctx->s_main));
    if (!flat.empty()) CK(cudaMemcpyAsync(sc.segs, flat.data(), flat.size() * sizeof(tb::Segment), cudaMemcpyHostToDevice, ctx->s_main));
    CK(cudaMemcpyAsync(sc.first, first.data(), first.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->s_main));
    CK(cudaStreamSynchronize(ctx->s_main));  // the host vectors die here; any stream may launch with the schedule afterwards
    *out = &(ctx->tb2_sched[{xb, xe}] = sc);
    return WAFER_OK;
}

// two lattice steps: psi[src] -> psi[src^1] for planes [xb, xe).  peer_lo / peer_hi != nullptr: the output planes
// [xb, lo_end) / [hi_begin, xe) are also stored into the lower / upper neighbour's buffer, at plane + shift_lo / shift_hi.
struct PeerTargets {
    double* lo = nullptr;
    double* hi = nullptr;
    long long shift_lo = 0, shift_hi = 0;
    int lo_end = INT_MIN, hi_begin = INT_MAX;
};
// The time-tiled sweep can read h = (dt*v)/2 instead of V: same bytes, two DMULs less per site and sweep pair, at the price
// of one more field in HBM.  Built lazily on the main stream; every launch of the sweep is ordered after it there.
// The one-step TMA sweeps can read h too (WAFER_T1_HFIELD=1); measured: no difference for them (5/7-point, excited
// steps within 0.5 %, gpurun_out/r2z) — they are not limited by those two multiplications — so it is off by default.
bool t1_reads_h() {
    static const bool on = getenv("WAFER_T1_HFIELD") && atoi(getenv("WAFER_T1_HFIELD")) == 1;
    return on;
}

int ensure_hfield(wafer_ctx* ctx) {
    static const bool want = !(getenv("WAFER_TB_HFIELD") && atoi(getenv("WAFER_TB_HFIELD")) == 0);
    if (!want || ctx->h_valid) return WAFER_OK;
    if (!ctx->hfield) {
        if (cudaMalloc(&ctx->hfield, ctx->bytes()) != cudaSuccess) { cudaGetLastError(); ctx->hfield = nullptr; return WAFER_OK; }  // no room: keep V
        TRY(make_tensor_map(ctx, &ctx->tm_h, ctx->hfield, tb::R1));
        if (ctx->use_t1) {
            const int rows = ctx->p.ext == 1 ? t1::Cfg<1>::TY : (ctx->p.ext == 2 ? t1::Cfg<2>::TY : t1::Cfg<3>::TY);
            TRY(make_tensor_map(ctx, &ctx->t1_h, ctx->hfield, rows));
        }
    }
    const long long n = ctx->g.total();
    build_h_kernel<<<ew_grid(ctx, n), EW_THREADS, 0, ctx->s_main>>>(ctx->v, ctx->hfield, n, ctx->p.dt);
    TRY(post_launch(ctx));
    ctx->h_valid = true;
    return WAFER_OK;
}

int launch_sweep_tb2(wafer_ctx* ctx, int src, int xb, int xe, cudaStream_t st, const PeerTargets* pt = nullptr) {
    if (xe <= xb) return WAFER_OK;
    const Geom& g = ctx->g;
    const bool hf = ctx->h_valid;
    const Tb2Sched* sc = nullptr;
    TRY(tb2_schedule(ctx, xb, xe, &sc));
    double* out = ctx->psi[src ^ 1];
    const tb::Sched ks{sc->bulk, sc->nbulk, sc->segs, sc->first, sc->counters};
    if (pt && (pt->lo || pt->hi)) {
        tb::PeerStores pr{0, 0, INT_MIN, INT_MAX};
        if (pt->lo) { pr.delta_lo = (pt->lo - out) + pt->shift_lo * g.plane; pr.lo_end = pt->lo_end; }
        if (pt->hi) { pr.delta_hi = (pt->hi - out) + pt->shift_hi * g.plane; pr.hi_begin = pt->hi_begin; }
        if (hf) tb::sweep_tb2_kernel<true, true><<<sc->ncta, tb::THREADS, tb::SMEM_BYTES, st>>>(ctx->tm_psi[src], ctx->tm_h, out, pr, g, ks, ctx->p.dt, denominator(ctx), ctx->den_ok);
        else tb::sweep_tb2_kernel<true, false><<<sc->ncta, tb::THREADS, tb::SMEM_BYTES, st>>>(ctx->tm_psi[src], ctx->tm_v, out, pr, g, ks, ctx->p.dt, denominator(ctx), ctx->den_ok);
    } else {
        const tb::PeerStores none{};
        if (hf) tb::sweep_tb2_kernel<false, true><<<sc->ncta, tb::THREADS, tb::SMEM_BYTES, st>>>(ctx->tm_psi[src], ctx->tm_h, out, none, g, ks, ctx->p.dt, denominator(ctx), ctx->den_ok);
        else tb::sweep_tb2_kernel<false, false><<<sc->ncta, tb::THREADS, tb::SMEM_BYTES, st>>>(ctx->tm_psi[src], ctx->tm_v, out, none, g, ks, ctx->p.dt, denominator(ctx), ctx->den_ok);
    }
    return post_launch(ctx);
}

// ---- TMA-pipelined one-step sweep (sweep_tma1.cuh), opt-in ---------------------------------------------------
template <int E>
int t1_tiles(const Geom& g) { return ceil_div(g.nz, t1::Cfg<E>::TZ) * ceil_div(g.ny, t1::Cfg<E>::TY); }

template <int E>
int t1_chunks(const wafer_ctx* ctx, int planes) {
    const long long tiles = t1_tiles<E>(ctx->g);
    int best_nc = 1;
    double best_cost = 1e300;
    for (int nc = 1; nc <= std::min(planes, 64); ++nc) {
        const double cost = (double)ceil_div(tiles * nc, ctx->sm_count) * ((double)ceil_div(planes, nc) + 2.0 * E);
        if (cost < best_cost * 0.999) { best_cost = cost; best_nc = nc; }
    }
    return best_nc;
}

// every instantiated mode of the TMA one-step kernel (sweep_tma1.cuh): sweeps with 0..5 running sums, observables with
// pot_sub none / scalar / array / energy only, sweeps that also leave the check's point-wise sums
#define T1_MODES(X) X(0) X(1) X(2) X(3) X(4) X(5) X(8) X(9) X(10) X(11) X(16) X(17) X(18)

template <int E>
int init_t1_e(wafer_ctx* ctx) {
    TRY(make_tensor_map(ctx, &ctx->t1_psi[0], ctx->psi[0], t1::Cfg<E>::R0));
    TRY(make_tensor_map(ctx, &ctx->t1_psi[1], ctx->psi[1], t1::Cfg<E>::R0));
    TRY(make_tensor_map(ctx, &ctx->t1_v, ctx->v, t1::Cfg<E>::TY));
    const int smem = (int)sizeof(t1::Smem<E>);
#define T1_ATTR(M) CK(cudaFuncSetAttribute(t1::sweep_tma1_kernel<E, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    T1_MODES(T1_ATTR)
#undef T1_ATTR
    return WAFER_OK;
}

int init_t1(wafer_ctx* ctx) {
    ctx->use_t1 = false;
    // the default one-step kernel whenever V is kept on the fly (13-24 % faster than the register-queue kernel for
    // 5/7-point, profiles/r1_tma1_one_step.md); boundary / interior sub-range launches of slab runs use it too
    const bool asked = (ctx->p.flags & WAFER_FLAG_TMA_ONE_STEP) != 0;
    const bool by_default = !(ctx->p.flags & WAFER_FLAG_SIMPLE_SWEEP);
    if (!ctx->onfly || !(asked || by_default)) return WAFER_OK;
    if (ctx->p.ext == 1) TRY(init_t1_e<1>(ctx));
    else if (ctx->p.ext == 2) TRY(init_t1_e<2>(ctx));
    else TRY(init_t1_e<3>(ctx));
    const double den = denominator(ctx);
    ctx->den_ok = (den > 7.888609052210118e-31 && den < 1.2676506002282294e30) ? 1 : 0;
    ctx->use_t1 = true;
    return WAFER_OK;
}

// mode (sweep_tma1.cuh): 0 plain sweep; 1 fused sum psi'^2; 1 + k: additionally the overlaps with stored states
// 0..k-1 (k <= 4); MODE_OBS + p: observables of psi[src], nothing stored; MODE_CHK + p: sweep + the check's point-wise sums.
// nb_total / bid_off: several launches (boundary + interior planes) share one row of per-CTA partial sums.
template <int E>
int launch_sweep_t1(wafer_ctx* ctx, int src, int xb, int xe, int mode, cudaStream_t st, int nb_total, int bid_off) {
    if (xe <= xb) return WAFER_OK;
    const Geom& g = ctx->g;
    const int planes = xe - xb, chunk = ceil_div(planes, t1_chunks<E>(ctx, planes));
    dim3 grid(ceil_div(g.nz, t1::Cfg<E>::TZ), ceil_div(g.ny, t1::Cfg<E>::TY), ceil_div(planes, chunk));
    const size_t smem = sizeof(t1::Smem<E>);
    t1::Extra ex{};
    if (mode < t1::MODE_OBS)
        for (int i = 0; i + 1 < mode; ++i) ex.q[i] = ctx->lowers[i];
    ex.potsub_arr = ctx->potsub_arr;
    ex.potsub = ctx->potsub_scalar;
    ex.nb_total = nb_total;
    ex.bid_off = bid_off;
    const bool sweep_mode = mode < t1::MODE_OBS || mode >= t1::MODE_CHK;
    ex.hf = t1_reads_h() && sweep_mode && ctx->h_valid ? 1 : 0;
    const CUtensorMap& tmv = ex.hf ? ctx->t1_h : ctx->t1_v;
    // the observables modes read psi[src] and store nothing: `out` only anchors the element offsets
    double* out = (mode >= t1::MODE_OBS && mode < t1::MODE_CHK) ? ctx->psi[src] : ctx->psi[src ^ 1];
#define T1_CASE(M)                                                                                                              \
    case M:                                                                                                                     \
        t1::sweep_tma1_kernel<E, M><<<grid, t1::Cfg<E>::THREADS, smem, st>>>(ctx->t1_psi[src], tmv, out, g, xb, xe, chunk,           \
                                                                             ctx->p.dt, denominator(ctx), ctx->den_ok, ctx->partials, ex); \
        break;
    switch (mode) {
        T1_MODES(T1_CASE)
        default: ctx->err = "internal: unknown mode of the TMA one-step kernel"; return WAFER_ERR_INVALID;
    }
#undef T1_CASE
    return post_launch(ctx);
}

template <int E>
int t1_blocks(const wafer_ctx* ctx, int xb, int xe) {
    const int planes = xe - xb, chunk = ceil_div(planes, t1_chunks<E>(ctx, planes));
    return t1_tiles<E>(ctx->g) * ceil_div(planes, chunk);
}

int sweep_blocks(const wafer_ctx* ctx, int xb, int xe) {
    if (!ctx->use_t1) return simple_blocks(ctx, xb, xe);
    return ctx->p.ext == 1 ? t1_blocks<1>(ctx, xb, xe) : (ctx->p.ext == 2 ? t1_blocks<2>(ctx, xb, xe) : t1_blocks<3>(ctx, xb, xe));
}

// one lattice step cur -> nxt for planes [xb, xe): TMA-pipelined kernel when enabled, register-queue kernel otherwise.
// nred > 1 (fused overlaps) is only available in the TMA kernel: callers ask fused_lowers() first.
int launch_sweep(wafer_ctx* ctx, const double* cur, double* nxt, int xb, int xe, int nred, long long part_off,
                 cudaStream_t st, int nb_total, int bid_off) {
    if (!ctx->use_t1) return launch_sweep_plain(ctx, cur, nxt, xb, xe, nred > 0, part_off, st);
    const int src = cur == ctx->psi[0] ? 0 : 1;
    if (nxt != ctx->psi[src ^ 1]) { ctx->err = "internal: sweep buffers are not the ping-pong pair"; return WAFER_ERR_INVALID; }
    switch (ctx->p.ext) {
        case 1: return launch_sweep_t1<1>(ctx, src, xb, xe, nred, st, nb_total, bid_off);
        case 2: return launch_sweep_t1<2>(ctx, src, xb, xe, nred, st, nb_total, bid_off);
        default: return launch_sweep_t1<3>(ctx, src, xb, xe, nred, st, nb_total, bid_off);
    }
}

// ---- inter-GPU ordering for the fused halo: monotone pass counters in peer-visible memory ------------------
// flags[0] is written by rank-1, flags[1] by rank+1 ("my boundary planes of pass n are in your ghost planes and I
// no longer read the ghost planes you are about to overwrite").  Spins are bounded (~30 s) so a dead peer cannot
// hang the GPU.
__global__ void p2p_wait_kernel(const volatile unsigned long long* flags, unsigned long long need, int has_lo, int has_hi,
                                int* timeout_flag) {
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
    for (int n = 0; n < 2; ++n) {
        if (!(n == 0 ? has_lo : has_hi)) continue;
        while (flags[n] < need) {
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
            if (t1 - t0 > 30000000000ull) { *timeout_flag = 1; return; }
            __nanosleep(200);
        }
    }
    __threadfence_system();  // acquire: the neighbour's ghost-plane stores precede its flag store
}
__global__ void p2p_signal_kernel(volatile unsigned long long* peer_lo, volatile unsigned long long* peer_hi,
                                  unsigned long long value) {
    __threadfence_system();  // release: every store of the preceding kernels in this stream (incl. peer stores)
    if (peer_lo) peer_lo[1] = value;  // I am rank-1's upper neighbour
    if (peer_hi) peer_hi[0] = value;  // I am rank+1's lower neighbour
    __threadfence_system();
}

// ghost-plane halo exchange of `buf` with the x neighbours (NCCL send/recv over NVLink)
int exchange(wafer_ctx* ctx, double* buf, cudaStream_t st) {
    if (ctx->world <= 1) return WAFER_OK;
    const Geom& g = ctx->g;
    // gx planes each way (gx = 2 for the time-tiled ThreePoint sweep, else ext): whole planes are contiguous
    const size_t cnt = (size_t)g.gx * g.plane;
    NcclApi& n = nccl_api();
    NK(n.GroupStart());
    if (ctx->rank > 0) {
        NK(n.Send(buf + g.off(0, -g.e, 0), cnt, kNcclFloat64, ctx->rank - 1, ctx->comm, st));
        NK(n.Recv(buf + g.off(-g.gx, -g.e, 0), cnt, kNcclFloat64, ctx->rank - 1, ctx->comm, st));
    }
    if (ctx->rank < ctx->world - 1) {
        NK(n.Send(buf + g.off(g.L - g.gx, -g.e, 0), cnt, kNcclFloat64, ctx->rank + 1, ctx->comm, st));
        NK(n.Recv(buf + g.off(g.L, -g.e, 0), cnt, kNcclFloat64, ctx->rank + 1, ctx->comm, st));
    }
    NK(n.GroupEnd());
    return WAFER_OK;
}

// ---- normalise + Gram-Schmidt (grid.rs:465-492) in two passes: see kernels.cuh (dots / coefficients / projection) ----
LowerPtrs lower_group(const wafer_ctx* ctx, int first, int count) {
    LowerPtrs lw{};
    for (int i = 0; i < count; ++i) lw.q[i] = ctx->lowers[first + i];
    return lw;
}

// raw overlaps of `psi` with stored states [first, first+count), count <= GS_GROUP -> scal[SL_RAW + 1 + first ...]
int dots_group(wafer_ctx* ctx, const double* psi, int first, int count) {
    const Geom& g = ctx->g;
    const long long ob = (long long)g.gx * g.plane / 2, oe = (long long)(g.gx + g.L) * g.plane / 2;
    const int grid = ew_grid(ctx, oe - ob);
    const LowerPtrs lw = lower_group(ctx, first, count);
    switch (count) {
        case 1: dots_kernel<1><<<grid, EW_THREADS, 0, ctx->s_main>>>(psi, lw, ob, oe, ctx->partials); break;
        case 2: dots_kernel<2><<<grid, EW_THREADS, 0, ctx->s_main>>>(psi, lw, ob, oe, ctx->partials); break;
        case 3: dots_kernel<3><<<grid, EW_THREADS, 0, ctx->s_main>>>(psi, lw, ob, oe, ctx->partials); break;
        default: dots_kernel<4><<<grid, EW_THREADS, 0, ctx->s_main>>>(psi, lw, ob, oe, ctx->partials); break;
    }
    TRY(post_launch(ctx));
    gs_coeff_kernel<true, false><<<1, 1024, 0, ctx->s_main>>>(ctx->partials, grid, count, ctx->scal + SL_RAW + 1 + first, 0, nullptr,
                                                                nullptr, nullptr, 0, nullptr);
    return post_launch(ctx);
}

int project(wafer_ctx* ctx, double* psi, const double* norm2_ptr, int wnum) {
    const long long n2 = ctx->g.total() / 2;
    const int grid = ew_grid(ctx, n2);
    if (wnum == 0) {
        if (norm2_ptr) {
            project_kernel<0, true><<<grid, EW_THREADS, 0, ctx->s_main>>>(psi, n2, norm2_ptr, LowerPtrs{}, nullptr);
            TRY(post_launch(ctx));
        }
        return WAFER_OK;
    }
    for (int first = 0; first < wnum; first += GS_GROUP) {
        const int count = std::min(GS_GROUP, wnum - first);
        const LowerPtrs lw = lower_group(ctx, first, count);
        const double* coef = ctx->scal + SL_COEF + first;
        const bool nrm = first == 0 && norm2_ptr;
#define PROJ(K)                                                                                                  \
    if (nrm) project_kernel<K, true><<<grid, EW_THREADS, 0, ctx->s_main>>>(psi, n2, norm2_ptr, lw, coef);        \
    else project_kernel<K, false><<<grid, EW_THREADS, 0, ctx->s_main>>>(psi, n2, nullptr, lw, coef)
        switch (count) {
            case 1: PROJ(1); break;
            case 2: PROJ(2); break;
            case 3: PROJ(3); break;
            default: PROJ(4); break;
        }
#undef PROJ
        TRY(post_launch(ctx));
    }
    return WAFER_OK;
}

// psi <- normalise(psi, *norm2_ptr) (norm2_ptr == nullptr: skip) then orthogonalise against stored states [0, wnum).
// have: the raw overlaps with the first `have` states (and, in raw[0], sum psi^2) already sit in ctx->partials as
// rows 0..have of `sweep_nb` per-CTA partial sums, left there by the fused sweep.
int gs_apply(wafer_ctx* ctx, double* psi, const double* norm2_ptr, int wnum, int have_rows = 0, int sweep_nb = 0) {
    double* raw = ctx->scal + SL_RAW;
    const bool one_launch = ctx->world == 1 && have_rows == wnum + 1;  // everything came out of the sweep: reduce + solve at once
    if (one_launch) {
        gs_coeff_kernel<true, true><<<1, 1024, 0, ctx->s_main>>>(ctx->partials, sweep_nb, have_rows, raw, wnum, raw + 1, norm2_ptr,
                                                                   ctx->gram, GRAM_LD, ctx->scal + SL_COEF);
        TRY(post_launch(ctx));
        return project(ctx, psi, norm2_ptr, wnum);
    }
    if (have_rows > 0) {
        gs_coeff_kernel<true, false><<<1, 1024, 0, ctx->s_main>>>(ctx->partials, sweep_nb, have_rows, raw, 0, nullptr, nullptr, nullptr,
                                                                    0, nullptr);
        TRY(post_launch(ctx));
    }
    for (int first = std::max(have_rows - 1, 0); first < wnum; first += GS_GROUP)
        TRY(dots_group(ctx, psi, first, std::min(GS_GROUP, wnum - first)));
    if (ctx->world > 1) {
        // one all-reduce per step: [sum psi^2 (when the sweep produced it), overlaps 0..wnum)
        const int off = have_rows > 0 ? 0 : 1, cnt = wnum + 1 - off;
        if (cnt > 0) NK(nccl_api().AllReduce(raw + off, raw + off, cnt, kNcclFloat64, kNcclSum, ctx->comm, ctx->s_main));
    }
    if (wnum > 0) {
        gs_coeff_kernel<false, true><<<1, 32, 0, ctx->s_main>>>(nullptr, 0, 0, raw, wnum, raw + 1, norm2_ptr, ctx->gram, GRAM_LD,
                                                                  ctx->scal + SL_COEF);
        TRY(post_launch(ctx));
    }
    return project(ctx, psi, norm2_ptr, wnum);
}

// a state joins the store: measure its overlaps with the states already there (row of the Gram matrix)
int register_lower(wafer_ctx* ctx, double* q) {
    if (!ctx->gram) CK(cudaMalloc(&ctx->gram, (size_t)GRAM_LD * GRAM_LD * sizeof(double)));
    const int m = (int)ctx->lowers.size();
    for (int first = 0; first < m; first += GS_GROUP) TRY(dots_group(ctx, q, first, std::min(GS_GROUP, m - first)));
    if (m > 0) {
        double* raw = ctx->scal + SL_RAW + 1;
        if (ctx->world > 1) NK(nccl_api().AllReduce(raw, raw, m, kNcclFloat64, kNcclSum, ctx->comm, ctx->s_main));
        CK(cudaMemcpyAsync(ctx->gram + (size_t)m * GRAM_LD, raw, m * sizeof(double), cudaMemcpyDeviceToDevice, ctx->s_main));
    }
    ctx->lowers.push_back(q);
    return WAFER_OK;
}

int observables_device(wafer_ctx* ctx) {
    const Geom& g = ctx->g;
    if (ctx->use_t1) {
        // TMA-pipelined pass (sweep_tma1.cuh MODE_OBS): psi and V stream through the same shared-memory ring as the sweep
        const int nb = sweep_blocks(ctx, 0, g.L);
        if (ctx->chk_valid) {
            // norm2, <pot_sub> and <r2> rode on the last sweep of evolve (grid.rs:405-437 are point-wise in psi): only the
            // energy, which needs the neighbours of the final psi, is left
            TRY(launch_sweep(ctx, ctx->psi[ctx->cur], ctx->psi[ctx->cur ^ 1], 0, g.L, t1::MODE_OBS + 3, 0, ctx->s_main));
            gs_coeff_kernel<true, false><<<1, 1024, 0, ctx->s_main>>>(ctx->partials, nb, 1, ctx->scal + SL_OBS, 0, nullptr, nullptr,
                                                                        nullptr, 0, nullptr);
            TRY(post_launch(ctx));
            CK(cudaMemcpyAsync(ctx->scal + SL_OBS + 1, ctx->scal + SL_CHK, 3 * sizeof(double), cudaMemcpyDeviceToDevice, ctx->s_main));
            if (ctx->world > 1)
                NK(nccl_api().AllReduce(ctx->scal + SL_OBS, ctx->scal + SL_OBS, 4, kNcclFloat64, kNcclSum, ctx->comm, ctx->s_main));
            return WAFER_OK;
        }
        TRY(launch_sweep(ctx, ctx->psi[ctx->cur], ctx->psi[ctx->cur ^ 1], 0, g.L, t1::MODE_OBS + ctx->potsub_mode, 0, ctx->s_main));
        return finalize(ctx, 4, nb, SL_OBS, ctx->s_main);
    }
    dim3 grid(ceil_div(g.nz, SW_BX), ceil_div(g.ny, SW_BY), ceil_div(g.L, SW_XCH)), block(SW_BX, SW_BY);
    const int nb = grid.x * grid.y * grid.z;
    const double den = denominator(ctx);
    const double* cur = ctx->psi[ctx->cur];
#define OBS(E, M) observables_kernel<E, M><<<grid, block, 0, ctx->s_main>>>(cur, ctx->v, ctx->potsub_arr, ctx->potsub_scalar, g, den, ctx->partials)
#define OBS_E(E)                                   \
    if (ctx->potsub_mode == 0) OBS(E, 0);          \
    else if (ctx->potsub_mode == 1) OBS(E, 1);     \
    else OBS(E, 2);
    if (ctx->p.ext == 1) { OBS_E(1) } else if (ctx->p.ext == 2) { OBS_E(2) } else { OBS_E(3) }
#undef OBS_E
#undef OBS
    TRY(post_launch(ctx));
    return finalize(ctx, 4, nb, SL_OBS, ctx->s_main);
}

// a fused-halo wait that gave up (dead or desynchronised neighbour) must not go unnoticed: every call that hands
// numbers back to the host checks the device-side timeout flag after its synchronisation
int check_p2p_timeout(wafer_ctx* ctx) {
    if (!ctx->p2p_timeout) return WAFER_OK;
    int t = 0;
    CK(cudaMemcpyAsync(&t, ctx->p2p_timeout, sizeof(int), cudaMemcpyDeviceToHost, ctx->s_main));
    CK(cudaStreamSynchronize(ctx->s_main));
    if (t) { ctx->err = "fused halo: timed out waiting for a neighbour GPU's pass flag"; return WAFER_ERR_NCCL; }
    return WAFER_OK;
}

int read_scalars(wafer_ctx* ctx, int slot, int n, double* out) {
    CK(cudaMemcpyAsync(ctx->h_scal + slot, ctx->scal + slot, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->s_main));
    CK(cudaStreamSynchronize(ctx->s_main));
    for (int i = 0; i < n; ++i) out[i] = ctx->h_scal[slot + i];
    return check_p2p_timeout(ctx);
}

// ---- host <-> device layout ------------------------------------------------------------------------------
// planes of the reference's padded array held by this rank (owned + ghosts), clipped to the array
void chunk_range(const wafer_ctx* ctx, bool worksized, long long* hp0, long long* hp1) {
    const Geom& g = ctx->g;
    const long long off = worksized ? 0 : g.e, top = worksized ? g.gnx : g.gnx + 2 * g.e;
    *hp0 = std::max<long long>(0, g.x0 + off - g.gx);
    *hp1 = std::min<long long>(top, g.x0 + g.L + off + g.gx);
}

void owned_range(const wafer_ctx* ctx, long long* hp0, long long* hp1) {
    // planes this rank is responsible for in the caller's global array: owned, plus the outer ring at the ends
    const Geom& g = ctx->g;
    *hp0 = g.x0 + g.e;
    *hp1 = g.x0 + g.L + g.e;
    if (ctx->rank == 0) *hp0 = 0;
    if (ctx->rank == ctx->world - 1) *hp1 = g.gnx + 2 * g.e;
}

// Bounce buffers: 2 x ~128 MB regardless of the lattice size (a 1024^3 field is 8.8 GB, a full-size staging copy
// of it was round 1's largest avoidable allocation).
int ensure_staging(wafer_ctx* ctx) {
    if (ctx->stg[0]) return WAFER_OK;
    const Geom& g = ctx->g;
    const long long plane_h = (long long)(g.ny + 2 * g.e) * (g.nz + 2 * g.e);
    const long long want = std::max<long long>(1, (128ll << 20) / (long long)(plane_h * sizeof(double)));
    ctx->stg_planes = std::min<long long>(want, g.L + 2 * g.gx + 2 * g.e);
    for (int b = 0; b < 2; ++b) {
        CK(cudaMalloc(&ctx->stg[b], (size_t)(ctx->stg_planes * plane_h) * sizeof(double)));
        CK(cudaEventCreateWithFlags(&ctx->ev_stg_free[b], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&ctx->ev_stg_full[b], cudaEventDisableTiming));
    }
    CK(cudaStreamCreateWithFlags(&ctx->s_copy, cudaStreamNonBlocking));
    return WAFER_OK;
}

// host -> device.  `host` points at plane `host_p0` of the reference array; planes [hp0, hp1) of it are read.
// Device planes whose host plane lies outside [hp0, hp1) are zeroed (ghost planes a later halo exchange fills).
int upload(wafer_ctx* ctx, const double* host, long long host_p0, long long hp0, long long hp1, double* dst, bool worksized,
           bool check_ring) {
    const Geom& g = ctx->g;
    const long long py = worksized ? g.ny : g.ny + 2 * g.e, pz = worksized ? g.nz : g.nz + 2 * g.e;
    const int hoff = worksized ? 0 : g.e;
    TRY(ensure_staging(ctx));
    if (check_ring) CK(cudaMemsetAsync(ctx->ring_flag, 0, sizeof(int), ctx->s_main));
    CK(cudaEventRecord(ctx->ev_stg_free[0], ctx->s_main));  // everything queued so far precedes the first refill
    CK(cudaEventRecord(ctx->ev_stg_free[1], ctx->s_main));
    int k = 0;
    for (int i0 = -g.gx; i0 < g.L + g.gx; i0 += (int)ctx->stg_planes, ++k) {
        const int i1 = (int)std::min<long long>(i0 + ctx->stg_planes, g.L + g.gx), b = k & 1;
        // host planes behind device planes [i0, i1), clipped to what the caller handed over
        const long long sp0 = std::max<long long>(hp0, g.x0 + i0 + hoff), sp1 = std::min<long long>(hp1, g.x0 + i1 + hoff);
        if (sp1 > sp0) {
            CK(cudaStreamWaitEvent(ctx->s_copy, ctx->ev_stg_free[b], 0));
            CK(cudaMemcpyAsync(ctx->stg[b], host + (sp0 - host_p0) * py * pz, (size_t)((sp1 - sp0) * py * pz) * sizeof(double),
                               cudaMemcpyHostToDevice, ctx->s_copy));
            CK(cudaEventRecord(ctx->ev_stg_full[b], ctx->s_copy));
            CK(cudaStreamWaitEvent(ctx->s_main, ctx->ev_stg_full[b], 0));
        }
        const int grid = (int)std::min<long long>((long long)(i1 - i0) * g.yp, (long long)ctx->sm_count * 16);
        if (check_ring) unpack_kernel<true><<<grid, 256, 0, ctx->s_main>>>(ctx->stg[b], dst, g, i0, i1, sp0, std::max(sp0, sp1), worksized ? 1 : 0, ctx->ring_flag);
        else unpack_kernel<false><<<grid, 256, 0, ctx->s_main>>>(ctx->stg[b], dst, g, i0, i1, sp0, std::max(sp0, sp1), worksized ? 1 : 0, ctx->ring_flag);
        TRY(post_launch(ctx));
        CK(cudaEventRecord(ctx->ev_stg_free[b], ctx->s_main));
    }
    if (check_ring) {
        int flag = 0;
        CK(cudaMemcpyAsync(&flag, ctx->ring_flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->s_main));
        CK(cudaStreamSynchronize(ctx->s_main));
        if (flag) {
            ctx->err = "padding ring of the wavefunction is not zero (config.rs:597-622 guarantees it)";
            return WAFER_ERR_RING_NONZERO;
        }
    } else {
        CK(cudaStreamSynchronize(ctx->s_main));  // the caller may free its buffer on return
    }
    return WAFER_OK;
}

// device -> host: padded planes [hp0, hp1) of the global array into `host`, which points at plane host_p0
int download(wafer_ctx* ctx, const double* src, double* host, long long host_p0, long long hp0, long long hp1) {
    const Geom& g = ctx->g;
    const long long py = g.ny + 2 * g.e, pz = g.nz + 2 * g.e;
    TRY(ensure_staging(ctx));
    CK(cudaEventRecord(ctx->ev_stg_free[0], ctx->s_main));
    CK(cudaEventRecord(ctx->ev_stg_free[1], ctx->s_main));
    int k = 0;
    for (long long sp0 = hp0; sp0 < hp1; sp0 += ctx->stg_planes, ++k) {
        const long long sp1 = std::min(hp1, sp0 + ctx->stg_planes);
        const int b = k & 1;
        CK(cudaStreamWaitEvent(ctx->s_main, ctx->ev_stg_free[b], 0));
        const int grid = (int)std::min<long long>((sp1 - sp0) * py, (long long)ctx->sm_count * 16);
        pack_kernel<<<grid, 256, 0, ctx->s_main>>>(src, ctx->stg[b], g, sp0, sp1, 0);
        TRY(post_launch(ctx));
        CK(cudaEventRecord(ctx->ev_stg_full[b], ctx->s_main));
        CK(cudaStreamWaitEvent(ctx->s_copy, ctx->ev_stg_full[b], 0));
        CK(cudaMemcpyAsync(host + (sp0 - host_p0) * py * pz, ctx->stg[b], (size_t)((sp1 - sp0) * py * pz) * sizeof(double),
                           cudaMemcpyDeviceToHost, ctx->s_copy));
        CK(cudaEventRecord(ctx->ev_stg_free[b], ctx->s_copy));
    }
    CK(cudaStreamSynchronize(ctx->s_copy));
    CK(cudaStreamSynchronize(ctx->s_main));
    return WAFER_OK;
}

// global-array entry points: this rank's chunk (owned + ghost planes) in, its owned planes out
int upload_global(wafer_ctx* ctx, const double* host, double* dst, bool worksized, bool check_ring) {
    long long hp0, hp1;
    chunk_range(ctx, worksized, &hp0, &hp1);
    return upload(ctx, host, 0, hp0, hp1, dst, worksized, check_ring);
}
int download_global(wafer_ctx* ctx, const double* src, double* host) {
    long long hp0, hp1;
    owned_range(ctx, &hp0, &hp1);
    return download(ctx, src, host, 0, hp0, hp1);
}

int alloc_field(wafer_ctx* ctx, double** p) {
    CK(cudaMalloc(p, ctx->bytes()));
    CK(cudaMemsetAsync(*p, 0, ctx->bytes(), ctx->s_main));
    return WAFER_OK;
}

int ensure_ab(wafer_ctx* ctx) {
    if (ctx->onfly) return WAFER_OK;
    if (!ctx->a) TRY(alloc_field(ctx, &ctx->a));
    if (!ctx->b) TRY(alloc_field(ctx, &ctx->b));
    const long long n = ctx->g.total();
    build_ab_kernel<<<ew_grid(ctx, n), EW_THREADS, 0, ctx->s_main>>>(ctx->v, ctx->a, ctx->b, n, ctx->p.dt);
    return post_launch(ctx);
}

int create_impl(const wafer_params* params, wafer_ctx* ctx) {
    const wafer_params& p = *params;
    REQUIRE(p.nx > 0 && p.ny > 0 && p.nz > 0, "grid.size must be positive");
    REQUIRE(p.ext >= 1 && p.ext <= 3, "ext must be 1 (ThreePoint), 2 (FivePoint) or 3 (SevenPoint)");
    REQUIRE(p.nx < (1u << 30) && p.ny < (1u << 30) && p.nz < (1u << 30), "grid.size too large");
    REQUIRE(std::isfinite(p.dn) && std::isfinite(p.dt) && std::isfinite(p.mass) && p.dn > 0 && p.mass != 0,
            "dn, dt, mass must be finite (dn > 0, mass != 0)");
    ctx->p = p;
    ctx->p.nccl_id = nullptr;
    ctx->world = p.world == 0 ? 1 : (int)p.world;
    ctx->rank = (int)p.rank;
    REQUIRE(ctx->rank < ctx->world, "rank must be < world");
    ctx->onfly = !(p.flags & WAFER_FLAG_AB_ARRAYS);

    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        ctx->err = "no CUDA device visible: this library has no CPU fallback";
        return WAFER_ERR_NO_DEVICE;
    }
    int dev = p.device;
    if (dev < 0) {
        const char* lr = getenv("LOCAL_RANK");
        dev = lr ? atoi(lr) % ndev : 0;
    }
    if (dev >= ndev) { ctx->err = "device ordinal out of range"; return WAFER_ERR_NO_DEVICE; }
    ctx->dev = dev;
    CK(cudaSetDevice(dev));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, dev));
    if (prop.major != 10) {
        ctx->err = std::string("device '") + prop.name + "' is not sm_100 (Blackwell B200); this build targets sm_100a only";
        return WAFER_ERR_NO_DEVICE;
    }
    ctx->sm_count = prop.multiProcessorCount;

    // slab decomposition along x
    Geom& g = ctx->g;
    uint64_t sx0, sx1;
    wafer_slab_partition(p.nx, (uint32_t)ctx->world, (uint32_t)ctx->rank, &sx0, &sx1);
    g.L = (int)(sx1 - sx0);
    g.x0 = (long long)sx0;
    g.ny = (int)p.ny; g.nz = (int)p.nz; g.e = (int)p.ext;
    g.gx = g.e == 1 ? 2 : g.e;  // the time-tiled ThreePoint sweep advances two steps per halo exchange
    g.yp = g.ny + 2 * g.e;
    g.zp = (int)(((long long)g.nz + 2 * g.e + 15) / 16 * 16);
    g.plane = (long long)g.yp * g.zp;
    g.gnx = p.nx; g.gny = p.ny; g.gnz = p.nz;
    ctx->min_L = (int)(p.nx / (uint64_t)ctx->world);  // wafer_slab_partition: the last ranks own floor(nx/world) planes
    REQUIRE(ctx->world == 1 || ctx->min_L >= g.gx, "every rank needs at least as many planes as the ghost depth (2 for ThreePoint, else ext)");

    int lo, hi;
    CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CK(cudaStreamCreateWithPriority(&ctx->s_main, cudaStreamNonBlocking, lo));
    CK(cudaStreamCreateWithPriority(&ctx->s_halo, cudaStreamNonBlocking, hi));
    CK(cudaEventCreate(&ctx->ev_t0));
    CK(cudaEventCreate(&ctx->ev_t1));
    CK(cudaEventCreateWithFlags(&ctx->ev_main, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&ctx->ev_halo, cudaEventDisableTiming));

    TRY(alloc_field(ctx, &ctx->psi[0]));
    TRY(alloc_field(ctx, &ctx->psi[1]));
    TRY(alloc_field(ctx, &ctx->v));
    const long long nb = std::max<long long>(
        {(long long)simple_blocks(ctx, 0, g.L) + 4 * simple_blocks(ctx, 0, g.e),
         (long long)ceil_div(g.nz, 56) * ceil_div(g.ny, 32) * std::min(g.L, 64),  // TMA one-step kernel: tiles x chunks
         (long long)ceil_div(g.nz, SW_BX) * ceil_div(g.ny, SW_BY) * ceil_div(g.L, SW_XCH), (long long)ctx->sm_count * 8});
    ctx->partials_cap = nb * 8;  // up to 1 + 4 fused running sums per CTA (sweep_tma1.cuh NRED), 4 for the observables
    CK(cudaMalloc(&ctx->partials, ctx->partials_cap * sizeof(double)));
    CK(cudaMalloc(&ctx->scal, SL_COUNT * sizeof(double)));
    CK(cudaMemsetAsync(ctx->scal, 0, SL_COUNT * sizeof(double), ctx->s_main));
    CK(cudaMallocHost(&ctx->h_scal, SL_COUNT * sizeof(double)));
    CK(cudaMalloc(&ctx->ring_flag, sizeof(int)));
    TRY(init_tb(ctx));
    TRY(init_t1(ctx));

    if (ctx->world > 1) {
        if (!p.nccl_id) { ctx->err = "world > 1 needs the 128-byte nccl_id of rank 0"; return WAFER_ERR_INVALID; }
        if (const char* why = nccl_api().load()) { ctx->err = why; return WAFER_ERR_NCCL; }
        NcclUniqueId id;
        memcpy(id.internal, p.nccl_id, sizeof(id.internal));
        NK(nccl_api().CommInitRank(&ctx->comm, ctx->world, id, ctx->rank));
    }
    CK(cudaStreamSynchronize(ctx->s_main));
    return WAFER_OK;
}

}  // namespace

// =============================================================================================================
extern "C" {

const char* wafer_version(void) { return "wafer_b200 0.1 (sm_100a)"; }

int wafer_create(const wafer_params* params, wafer_ctx** out) {
    if (!params || !out) { g_create_error = "NULL argument"; return WAFER_ERR_INVALID; }
    *out = nullptr;
    wafer_ctx* ctx = new (std::nothrow) wafer_ctx();
    if (!ctx) { g_create_error = "out of host memory"; return WAFER_ERR_INVALID; }
    int rc;
    try {
        rc = create_impl(params, ctx);
    } catch (const std::exception& e) {
        ctx->err = std::string("wafer_create: ") + e.what();
        rc = WAFER_ERR_INVALID;
    }
    if (rc != WAFER_OK) {
        g_create_error = ctx->err;
        wafer_destroy(ctx);
        return rc;
    }
    *out = ctx;
    return WAFER_OK;
}

int wafer_destroy(wafer_ctx* ctx) {
    if (!ctx) return WAFER_OK;
    if (ctx->s_main) cudaStreamSynchronize(ctx->s_main);
    if (ctx->s_halo) cudaStreamSynchronize(ctx->s_halo);
    drop_step_graphs(ctx);
    if (ctx->comm) nccl_api().CommDestroy(ctx->comm);
    for (int n = 0; n < 2; ++n) {
        if (ctx->peer_psi[n][0]) cudaIpcCloseMemHandle(ctx->peer_psi[n][0]);
        if (ctx->peer_psi[n][1]) cudaIpcCloseMemHandle(ctx->peer_psi[n][1]);
        if (ctx->peer_flags[n]) cudaIpcCloseMemHandle(ctx->peer_flags[n]);
    }
    cudaFree(ctx->flags); cudaFree(ctx->p2p_timeout);
    for (double* q : ctx->lowers) cudaFree(q);
    cudaFree(ctx->hfield);
    cudaFree(ctx->psi[0]); cudaFree(ctx->psi[1]); cudaFree(ctx->v); cudaFree(ctx->a); cudaFree(ctx->b);
    for (int b = 0; b < 2; ++b) {
        cudaFree(ctx->stg[b]);
        if (ctx->ev_stg_free[b]) cudaEventDestroy(ctx->ev_stg_free[b]);
        if (ctx->ev_stg_full[b]) cudaEventDestroy(ctx->ev_stg_full[b]);
    }
    if (ctx->s_copy) cudaStreamDestroy(ctx->s_copy);
    cudaFree(ctx->d_cksum); cudaFree(ctx->gram);
    for (auto& kv : ctx->tb2_sched) {
        cudaFree(kv.second.bulk); cudaFree(kv.second.segs); cudaFree(kv.second.first); cudaFree(kv.second.counters);
    }
    cudaFree(ctx->potsub_arr); cudaFree(ctx->partials); cudaFree(ctx->scal); cudaFree(ctx->ring_flag);
    if (ctx->h_scal) cudaFreeHost(ctx->h_scal);
    if (ctx->ev_t0) cudaEventDestroy(ctx->ev_t0);
    if (ctx->ev_t1) cudaEventDestroy(ctx->ev_t1);
    if (ctx->ev_main) cudaEventDestroy(ctx->ev_main);
    if (ctx->ev_halo) cudaEventDestroy(ctx->ev_halo);
    if (ctx->s_main) cudaStreamDestroy(ctx->s_main);
    if (ctx->s_halo) cudaStreamDestroy(ctx->s_halo);
    delete ctx;
    return WAFER_OK;
}

const char* wafer_last_error(const wafer_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int wafer_nccl_unique_id(uint8_t out[128]) {
    if (!out) return WAFER_ERR_INVALID;
    if (const char* why = nccl_api().load()) { g_create_error = why; return WAFER_ERR_NCCL; }
    NcclUniqueId id;
    if (nccl_api().GetUniqueId(&id) != kNcclSuccess) { g_create_error = "ncclGetUniqueId failed"; return WAFER_ERR_NCCL; }
    memcpy(out, id.internal, 128);
    return WAFER_OK;
}

int wafer_slab_partition(uint64_t nx, uint32_t world, uint32_t rank, uint64_t* x0, uint64_t* x1) {
    if (world == 0) world = 1;
    if (rank >= world || !x0 || !x1) return WAFER_ERR_INVALID;
    const uint64_t base = nx / world, rem = nx % world;
    *x0 = rank * base + std::min<uint64_t>(rank, rem);
    *x1 = *x0 + base + (rank < rem ? 1 : 0);
    return WAFER_OK;
}

int wafer_tb2_plan(uint32_t ny, uint32_t nz, int32_t xb, int32_t xe, uint32_t slots, int32_t* segments, uint64_t cap, uint64_t* n) {
    if (!n || ny == 0 || nz == 0 || xe <= xb || slots == 0) return WAFER_ERR_INVALID;
    try {
        const Tb2Plan pl = tb2_plan((int)ny, (int)nz, xb, xe, (int)slots, "dynamic", 512);
        uint64_t k = 0;
        auto put = [&](int owner, const tb::Segment& sg) {
            if (segments && k < cap) {
                int32_t* o = segments + 5 * k;
                o[0] = owner; o[1] = sg.y0; o[2] = sg.z0; o[3] = sg.xa; o[4] = sg.xz;
            }
            ++k;
        };
        for (const auto& sg : pl.bulk) put(-1, sg);
        for (size_t c = 0; c < pl.per.size(); ++c)
            for (const auto& sg : pl.per[c]) put((int)c, sg);
        *n = k;
        return WAFER_OK;
    } catch (const std::exception&) {
        return WAFER_ERR_INVALID;
    }
}

int wafer_slab(const wafer_ctx* ctx, uint64_t* x0, uint64_t* x1) {
    if (!ctx) return WAFER_ERR_INVALID;
    if (x0) *x0 = (uint64_t)ctx->g.x0;
    if (x1) *x1 = (uint64_t)(ctx->g.x0 + ctx->g.L);
    return WAFER_OK;
}

// ---- state in / out ------------------------------------------------------------------------------------------
int wafer_set_potential(wafer_ctx* ctx, const double* v_padded) {
    if (!ctx) return WAFER_ERR_INVALID;
    ctx->chk_valid = false;  // psi (or what the check sums depend on) changes
    REQUIRE(v_padded, "v_padded is NULL");
    CK(cudaSetDevice(ctx->dev));
    ctx->h_valid = false;
    TRY(upload_global(ctx, v_padded, ctx->v, false, false));
    TRY(ensure_ab(ctx));
    ctx->have_v = true;
    return WAFER_OK;
}

int wafer_get_potential(wafer_ctx* ctx, double* v_padded) {
    if (!ctx) return WAFER_ERR_INVALID;
    REQUIRE(v_padded, "v_padded is NULL");
    if (!ctx->have_v) { ctx->err = "potential not set"; return WAFER_ERR_NOT_READY; }
    CK(cudaSetDevice(ctx->dev));
    return download_global(ctx, ctx->v, v_padded);
}

int wafer_set_pot_sub_scalar(wafer_ctx* ctx, double c) {
    if (!ctx) return WAFER_ERR_INVALID;
    ctx->chk_valid = false;  // psi (or what the check sums depend on) changes
    // potential.rs:148-152: (None, Some(c)) only when c > 0, else (None, None)
    ctx->potsub_mode = c > 0.0 ? 1 : 0;
    ctx->potsub_scalar = c > 0.0 ? c : 0.0;
    return WAFER_OK;
}

int wafer_set_pot_sub_array(wafer_ctx* ctx, const double* work) {
    if (!ctx) return WAFER_ERR_INVALID;
    ctx->chk_valid = false;  // psi (or what the check sums depend on) changes
    REQUIRE(work, "pot_sub array is NULL");
    CK(cudaSetDevice(ctx->dev));
    if (!ctx->potsub_arr) TRY(alloc_field(ctx, &ctx->potsub_arr));
    TRY(upload_global(ctx, work, ctx->potsub_arr, true, false));
    ctx->potsub_mode = 2;
    return WAFER_OK;
}

int wafer_set_phi(wafer_ctx* ctx, const double* phi_padded) {
    if (!ctx) return WAFER_ERR_INVALID;
    ctx->chk_valid = false;  // psi (or what the check sums depend on) changes
    REQUIRE(phi_padded, "phi_padded is NULL");
    CK(cudaSetDevice(ctx->dev));
    const int rc = upload_global(ctx, phi_padded, ctx->psi[ctx->cur], false, true);
    ctx->have_phi = rc == WAFER_OK;
    return rc;
}

int wafer_get_phi(wafer_ctx* ctx, double* phi_padded) {
    if (!ctx) return WAFER_ERR_INVALID;
    REQUIRE(phi_padded, "phi_padded is NULL");
    if (!ctx->have_phi) { ctx->err = "phi not set"; return WAFER_ERR_NOT_READY; }
    CK(cudaSetDevice(ctx->dev));
    return download_global(ctx, ctx->psi[ctx->cur], phi_padded);
}

int wafer_slab_planes(const wafer_ctx* ctx, int32_t which, uint64_t* p0, uint64_t* p1) {
    if (!ctx) return WAFER_ERR_INVALID;
    long long a, b;
    if (which == 0) chunk_range(ctx, false, &a, &b);
    else owned_range(ctx, &a, &b);
    if (p0) *p0 = (uint64_t)a;
    if (p1) *p1 = (uint64_t)b;
    return WAFER_OK;
}

int wafer_set_phi_slab(wafer_ctx* ctx, const double* chunk) {
    if (!ctx) return WAFER_ERR_INVALID;
    ctx->chk_valid = false;  // psi (or what the check sums depend on) changes
    REQUIRE(chunk, "chunk is NULL");
    CK(cudaSetDevice(ctx->dev));
    long long hp0, hp1;
    chunk_range(ctx, false, &hp0, &hp1);
    const int rc = upload(ctx, chunk, hp0, hp0, hp1, ctx->psi[ctx->cur], false, true);
    ctx->have_phi = rc == WAFER_OK;
    return rc;
}

int wafer_set_phi_owned(wafer_ctx* ctx, const double* chunk) {
    if (!ctx) return WAFER_ERR_INVALID;
    ctx->chk_valid = false;  // psi (or what the check sums depend on) changes
    REQUIRE(chunk, "chunk is NULL");
    CK(cudaSetDevice(ctx->dev));
    long long hp0, hp1;
    owned_range(ctx, &hp0, &hp1);
    const int rc = upload(ctx, chunk, hp0, hp0, hp1, ctx->psi[ctx->cur], false, true);
    ctx->have_phi = rc == WAFER_OK;
    if (rc != WAFER_OK) return rc;
    return exchange(ctx, ctx->psi[ctx->cur], ctx->s_main);  // ghost planes come from the neighbours (collective)
}

int wafer_get_phi_slab(wafer_ctx* ctx, double* chunk) {
    if (!ctx) return WAFER_ERR_INVALID;
    REQUIRE(chunk, "chunk is NULL");
    if (!ctx->have_phi) { ctx->err = "phi not set"; return WAFER_ERR_NOT_READY; }
    CK(cudaSetDevice(ctx->dev));
    long long hp0, hp1;
    owned_range(ctx, &hp0, &hp1);
    return download(ctx, ctx->psi[ctx->cur], chunk, hp0, hp0, hp1);
}

static int wafer_push_lower_impl(wafer_ctx* ctx, const double* q_padded) {
    if (!ctx) return WAFER_ERR_INVALID;
    REQUIRE(q_padded, "q_padded is NULL");
    REQUIRE(ctx->lowers.size() < 255, "at most 255 lower states (wavenum is a u8, config.rs:308)");
    CK(cudaSetDevice(ctx->dev));
    double* q = nullptr;
    TRY(alloc_field(ctx, &q));
    int rc = upload_global(ctx, q_padded, q, false, true);
    if (rc == WAFER_OK) rc = register_lower(ctx, q);
    if (rc != WAFER_OK) cudaFree(q);
    return rc;
}

static int wafer_push_lower_from_phi_impl(wafer_ctx* ctx) {
    if (!ctx) return WAFER_ERR_INVALID;
    if (!ctx->have_phi) { ctx->err = "phi not set"; return WAFER_ERR_NOT_READY; }
    REQUIRE(ctx->lowers.size() < 255, "at most 255 lower states (wavenum is a u8, config.rs:308)");
    CK(cudaSetDevice(ctx->dev));
    double* q = nullptr;
    CK(cudaMalloc(&q, ctx->bytes()));
    CK(cudaMemcpyAsync(q, ctx->psi[ctx->cur], ctx->bytes(), cudaMemcpyDeviceToDevice, ctx->s_main));
    const int rc = register_lower(ctx, q);
    if (rc != WAFER_OK) cudaFree(q);
    return rc;
}

int wafer_get_lower(wafer_ctx* ctx, uint32_t idx, double* q_padded) {
    if (!ctx) return WAFER_ERR_INVALID;
    REQUIRE(q_padded && idx < ctx->lowers.size(), "no such lower state");
    CK(cudaSetDevice(ctx->dev));
    return download_global(ctx, ctx->lowers[idx], q_padded);
}

int wafer_phi_from_lower(wafer_ctx* ctx, uint32_t idx) {
    if (!ctx) return WAFER_ERR_INVALID;
    ctx->chk_valid = false;  // psi (or what the check sums depend on) changes
    REQUIRE(idx < ctx->lowers.size(), "no such lower state");
    CK(cudaSetDevice(ctx->dev));
    CK(cudaMemcpyAsync(ctx->psi[ctx->cur], ctx->lowers[idx], ctx->bytes(), cudaMemcpyDeviceToDevice, ctx->s_main));
    ctx->have_phi = true;
    return WAFER_OK;
}

int wafer_phi_seed_from_lower(wafer_ctx* ctx, uint32_t idx) {
    if (!ctx) return WAFER_ERR_INVALID;
    ctx->chk_valid = false;  // psi (or what the check sums depend on) changes
    REQUIRE(idx < ctx->lowers.size(), "no such lower state");
    CK(cudaSetDevice(ctx->dev));
    const long long rows = (long long)(ctx->g.L + 2 * ctx->g.gx) * ctx->g.ny;
    seed_from_state_kernel<<<(int)std::min<long long>(rows, (long long)ctx->sm_count * 32), 128, 0, ctx->s_main>>>(
        ctx->psi[ctx->cur], ctx->lowers[idx], ctx->g);
    TRY(post_launch(ctx));
    ctx->have_phi = true;
    return WAFER_OK;
}

int wafer_clear_lowers(wafer_ctx* ctx) {
    if (!ctx) return WAFER_ERR_INVALID;
    CK(cudaSetDevice(ctx->dev));
    CK(cudaStreamSynchronize(ctx->s_main));
    for (double* q : ctx->lowers) cudaFree(q);
    ctx->lowers.clear();
    drop_step_graphs(ctx);  // they hold the freed pointers
    return WAFER_OK;
}

uint32_t wafer_num_lowers(const wafer_ctx* ctx) { return ctx ? (uint32_t)ctx->lowers.size() : 0; }

int wafer_generate_potential(wafer_ctx* ctx, int32_t kind, double sig) {
    if (!ctx) return WAFER_ERR_INVALID;
    ctx->chk_valid = false;  // psi (or what the check sums depend on) changes
    CK(cudaSetDevice(ctx->dev));
    GenParams gp = make_gen_params(ctx->p.dn, ctx->p.mass, sig);
    REQUIRE(potential_kind_supported(kind), "PotentialNotAvailable: no formula for this potential kind (potential.rs:315-317)");
    const long long rows = (long long)(ctx->g.L + 2 * ctx->g.gx) * ctx->g.ny;
    ctx->h_valid = false;
    gen_potential_kernel<<<(int)std::min<long long>(rows, (long long)ctx->sm_count * 32), 128, 0, ctx->s_main>>>(ctx->v, ctx->g, kind, gp);
    TRY(post_launch(ctx));
    TRY(ensure_ab(ctx));
    ctx->have_v = true;
    return WAFER_OK;
}

int wafer_generate_initial_condition(wafer_ctx* ctx, int32_t kind) {
    if (!ctx) return WAFER_ERR_INVALID;
    ctx->chk_valid = false;  // psi (or what the check sums depend on) changes
    REQUIRE(kind >= 2 && kind <= 4, "only Coulomb (2), Constant (3) and Boolean (4) can be generated on the device");
    CK(cudaSetDevice(ctx->dev));
    GenParams gp = make_gen_params(ctx->p.dn, ctx->p.mass, 0.0);
    const long long rows = (long long)(ctx->g.L + 2 * ctx->g.gx) * ctx->g.ny;
    gen_ic_kernel<<<(int)std::min<long long>(rows, (long long)ctx->sm_count * 32), 128, 0, ctx->s_main>>>(ctx->psi[ctx->cur], ctx->g, kind, gp);
    TRY(post_launch(ctx));
    ctx->have_phi = true;
    return WAFER_OK;
}

// ---- the hot path --------------------------------------------------------------------------------------------
static int ready(wafer_ctx* ctx, bool need_v) {
    if (!ctx->have_phi) { ctx->err = "phi not set (wafer_set_phi / wafer_generate_initial_condition)"; return WAFER_ERR_NOT_READY; }
    if (need_v && !ctx->have_v) { ctx->err = "potential not set (wafer_set_potential / wafer_generate_potential)"; return WAFER_ERR_NOT_READY; }
    return WAFER_OK;
}

int wafer_observables_compute(wafer_ctx* ctx, wafer_observables* out) {
    if (!ctx) return WAFER_ERR_INVALID;
    REQUIRE(out, "out is NULL");
    TRY(ready(ctx, true));
    CK(cudaSetDevice(ctx->dev));
    TRY(observables_device(ctx));
    double s[4];
    TRY(read_scalars(ctx, SL_OBS, 4, s));
    out->energy = s[0]; out->norm2 = s[1]; out->v_infinity = s[2]; out->r2 = s[3];
    return WAFER_OK;
}

int wafer_norm2(wafer_ctx* ctx, double* out) {
    if (!ctx) return WAFER_ERR_INVALID;
    REQUIRE(out, "out is NULL");
    TRY(ready(ctx, false));
    CK(cudaSetDevice(ctx->dev));
    const Geom& g = ctx->g;
    const long long ob = (long long)g.gx * g.plane / 2, oe = (long long)(g.gx + g.L) * g.plane / 2;
    const int grid = ew_grid(ctx, oe - ob);
    norm2_kernel<<<grid, EW_THREADS, 0, ctx->s_main>>>(ctx->psi[ctx->cur], ob, oe, ctx->partials);
    TRY(post_launch(ctx));
    TRY(finalize(ctx, 1, grid, SL_TMP, ctx->s_main));
    return read_scalars(ctx, SL_TMP, 1, out);
}

int wafer_normalise(wafer_ctx* ctx, double norm2) {
    if (!ctx) return WAFER_ERR_INVALID;
    ctx->chk_valid = false;  // psi (or what the check sums depend on) changes
    TRY(ready(ctx, false));
    CK(cudaSetDevice(ctx->dev));
    set_scalar_kernel<<<1, 1, 0, ctx->s_main>>>(ctx->scal + SL_TMP, norm2);
    TRY(post_launch(ctx));
    return gs_apply(ctx, ctx->psi[ctx->cur], ctx->scal + SL_TMP, 0);
}

int wafer_orthogonalise(wafer_ctx* ctx, uint8_t wnum) {
    if (!ctx) return WAFER_ERR_INVALID;
    ctx->chk_valid = false;  // psi (or what the check sums depend on) changes
    TRY(ready(ctx, false));
    // grid.rs:478: w_store.iter().take(wnum) — silently clamps to the stored count
    const int n = std::min<int>(wnum, (int)ctx->lowers.size());
    if (n == 0) return WAFER_OK;
    CK(cudaSetDevice(ctx->dev));
    return gs_apply(ctx, ctx->psi[ctx->cur], nullptr, n);
}

// one excited-state step on a single rank (grid.rs:567-681): sweep with the fused sums, then normalise + Gram-Schmidt
static int excited_step(wafer_ctx* ctx, int wnum) {
    const Geom& g = ctx->g;
    const int src = ctx->cur;
    const int nred = ctx->use_t1 ? 1 + std::min(wnum, t1::MAX_FUSED_LOWERS) : 1;
    TRY(launch_sweep(ctx, ctx->psi[src], ctx->psi[src ^ 1], 0, g.L, nred, 0, ctx->s_main));
    ctx->cur ^= 1;
    return gs_apply(ctx, ctx->psi[src ^ 1], ctx->scal + SL_RAW, wnum, nred, sweep_blocks(ctx, 0, g.L));
}


// `pairs` times two excited-state steps through a captured graph.  Returns WAFER_OK with *ran = false when graphs are not
// to be used (several ranks, switched off, or the capture failed: the caller then steps the ordinary way).
static int excited_pairs_graph(wafer_ctx* ctx, int wnum, uint64_t pairs, bool* ran) {
    *ran = false;
    static const bool on = !(getenv("WAFER_GRAPHS") && atoi(getenv("WAFER_GRAPHS")) == 0);
    if (!on || ctx->world != 1 || pairs == 0) return WAFER_OK;
    const std::pair<int, int> key{wnum, ctx->cur};
    auto it = ctx->step_graphs.find(key);
    if (it == ctx->step_graphs.end()) {
        cudaGraph_t graph = nullptr;
        cudaGraphExec_t exec = nullptr;
        const uint64_t launches0 = ctx->launches;
        const int cur0 = ctx->cur;
        if (cudaStreamBeginCapture(ctx->s_main, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); return WAFER_OK; }
        int rc = excited_step(ctx, wnum);
        if (rc == WAFER_OK) rc = excited_step(ctx, wnum);
        const cudaError_t ce = cudaStreamEndCapture(ctx->s_main, &graph);
        const int per_pair = (int)(ctx->launches - launches0);
        ctx->launches = launches0;  // nothing ran yet
        ctx->cur = cur0;
        if (rc != WAFER_OK || ce != cudaSuccess || !graph || cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess) {
            if (graph) cudaGraphDestroy(graph);
            cudaGetLastError();
            return WAFER_OK;
        }
        cudaGraphDestroy(graph);
        it = ctx->step_graphs.emplace(key, std::make_pair(exec, per_pair)).first;
    }
    for (uint64_t i = 0; i < pairs; ++i) CK(cudaGraphLaunch(it->second.first, ctx->s_main));
    ctx->launches += pairs * (uint64_t)it->second.second;
    *ran = true;
    return WAFER_OK;
}

static int wafer_evolve_impl(wafer_ctx* ctx, uint8_t wnum_in, uint64_t steps) {
    if (!ctx) return WAFER_ERR_INVALID;
    TRY(ready(ctx, true));
    CK(cudaSetDevice(ctx->dev));
    const Geom& g = ctx->g;
    const int wnum = std::min<int>(wnum_in, (int)ctx->lowers.size());
    const bool excited = wnum_in > 0;  // grid.rs:674: norm/normalise run for wnum > 0 even with an empty w_store
    // every rank must take the same branch (one side calling NCCL while the other stores through peer memory would
    // hang): decide from the smallest slab of the decomposition, not from this rank's own L
    const bool overlap = ctx->world > 1 && !excited && ctx->min_L > 2 * g.gx;
    // (before the events below: the halo stream's launches must come after the build of h as well)
    if (((ctx->use_tb && !excited) || (ctx->use_t1 && t1_reads_h())) && steps >= 4)
        TRY(ensure_hfield(ctx));  // one 16 B/site pass: pays for itself after two sweep pairs
    if (overlap) {
        CK(cudaEventRecord(ctx->ev_main, ctx->s_main));
        CK(cudaEventRecord(ctx->ev_halo, ctx->s_main));
    }
    const uint64_t total = steps == 0 ? 1 : steps;  // grid.rs:562-686 is a do-while: steps == 0 still sweeps once
    const bool fused = overlap && ctx->p2p;  // halo stores fused into the boundary kernels (peer memory), no NCCL
    // North-star "block reductions fused into the final sweep before each check": the LAST step of a ground-state evolve
    // runs as a one-step sweep that also accumulates sum psi^2, sum psi^2 pot_sub and sum psi^2 r2 of its output; the
    // check that follows then only needs the energy pass.  Worth one non-time-tiled pass only on long calls.
    const bool fuse_chk = ctx->use_t1 && !excited && total >= 64 && !(ctx->p.flags & WAFER_FLAG_NO_FUSED_CHECK);
    ctx->chk_valid = false;
    int chk_nb = 0;
    const int has_lo = ctx->rank > 0, has_hi = ctx->rank < ctx->world - 1;
    uint64_t done = 0;
    if (excited && ctx->world == 1 && wnum <= t1::MAX_FUSED_LOWERS && total >= 8) {
        bool ran = false;
        TRY(excited_pairs_graph(ctx, wnum, total / 2, &ran));
        if (ran) done = (total / 2) * 2;
    }
    while (done < total) {
        // ground state: two steps per HBM pass with the time-tiled TMA kernel whenever two steps remain;
        // excited states need a global norm / Gram-Schmidt after EVERY step (grid.rs:674-681): one step per pass
        const bool last = total - done == 1;
        const bool two = ctx->use_tb && !excited && total - done >= 2 && !(fuse_chk && total - done == 2);
        const int chk_mode = (fuse_chk && last) ? t1::MODE_CHK + ctx->potsub_mode : 0;
        const int b = g.gx;  // boundary planes = ghost depth the neighbours keep (2 for ThreePoint, else ext)
        const int src = ctx->cur;
        const double* cur = ctx->psi[src];
        double* nxt = ctx->psi[src ^ 1];
        // Fused halos, two-step pass, "whole column" form (default): ONE launch on the main stream sweeps all owned planes and
        // stores its first / last two output planes into the neighbours' ghost planes as it produces them; the neighbours'
        // pass counters are awaited before it and published after it.  No separate boundary launches: those produce 2 planes
        // for 6 plane-iterations of pipeline (140 instead of 132 iterations per tile column at 128 planes per rank, and four
        // small launches that the persistent interior kernel serialises behind itself anyway — 0.894 parallel efficiency at
        // N = 8, profiles/scaling_r2).  The price is a per-pass handshake between neighbours instead of a pass of slack.
        static const bool split_halo = getenv("WAFER_P2P_SPLIT") && atoi(getenv("WAFER_P2P_SPLIT")) == 1;
        if (fused && two && !split_halo) {
            CK(cudaStreamWaitEvent(ctx->s_main, ctx->ev_halo, 0));  // an earlier pass may have used the halo stream
            if (ctx->dbg_halo_delay_ns) {
                spin_kernel<<<1, 1, 0, ctx->s_main>>>(ctx->dbg_halo_delay_ns);
                TRY(post_launch(ctx));
            }
            // counter >= P: the neighbours have finished pass P-1 — their stores into my ghost planes of `cur` have landed,
            // and they no longer read the ghost planes (of their `nxt`) that this pass overwrites
            p2p_wait_kernel<<<1, 1, 0, ctx->s_main>>>(ctx->flags, ctx->pass, has_lo, has_hi, ctx->p2p_timeout);
            TRY(post_launch(ctx));
            PeerTargets pt;
            if (has_lo) { pt.lo = ctx->peer_psi[0][src ^ 1]; pt.shift_lo = ctx->peer_L[0]; pt.lo_end = b; }
            if (has_hi) { pt.hi = ctx->peer_psi[1][src ^ 1]; pt.shift_hi = -(long long)g.L; pt.hi_begin = g.L - b; }
            TRY(launch_sweep_tb2(ctx, src, 0, g.L, ctx->s_main, &pt));
            ctx->pass += 1;
            p2p_signal_kernel<<<1, 1, 0, ctx->s_main>>>(has_lo ? ctx->peer_flags[0] : nullptr, has_hi ? ctx->peer_flags[1] : nullptr, ctx->pass);
            TRY(post_launch(ctx));
            CK(cudaEventRecord(ctx->ev_main, ctx->s_main));
            CK(cudaEventRecord(ctx->ev_halo, ctx->s_main));
        } else if (overlap) {
            // boundary planes + halo on the high-priority stream, interior on the main stream
            CK(cudaStreamWaitEvent(ctx->s_halo, ctx->ev_main, 0));
            CK(cudaStreamWaitEvent(ctx->s_main, ctx->ev_halo, 0));
            if (ctx->dbg_halo_delay_ns) {
                spin_kernel<<<1, 1, 0, ctx->s_halo>>>(ctx->dbg_halo_delay_ns);
                TRY(post_launch(ctx));
            }
            if (fused) {
                // my ghost planes of `cur` hold the neighbours' pass-(n-1) boundary planes once their flag says so;
                // the same flag says they are done reading the ghost planes (of their `nxt`) that I overwrite now
                p2p_wait_kernel<<<1, 1, 0, ctx->s_halo>>>(ctx->flags, ctx->pass, has_lo, has_hi, ctx->p2p_timeout);
                TRY(post_launch(ctx));
            }
            double* plo = fused && has_lo ? ctx->peer_psi[0][src ^ 1] : nullptr;
            double* phi_ = fused && has_hi ? ctx->peer_psi[1][src ^ 1] : nullptr;
            if (two) {
                // my planes [0,b) are the lower neighbour's ghost planes [L_lo, L_lo+b); my [L-b,L) the upper one's [-b,0)
                PeerTargets lo_t, hi_t;
                if (plo) { lo_t.lo = plo; lo_t.shift_lo = ctx->peer_L[0]; lo_t.lo_end = b; }
                if (phi_) { hi_t.hi = phi_; hi_t.shift_hi = -(long long)g.L; hi_t.hi_begin = g.L - b; }
                TRY(launch_sweep_tb2(ctx, src, 0, b, ctx->s_halo, &lo_t));
                TRY(launch_sweep_tb2(ctx, src, g.L - b, g.L, ctx->s_halo, &hi_t));
            } else {
                // (final step with fused check sums: the three launches share one row of per-CTA partial sums)
                const int nb_lo = sweep_blocks(ctx, 0, b), nb_hi = sweep_blocks(ctx, g.L - b, g.L);
                chk_nb = chk_mode ? nb_lo + nb_hi + sweep_blocks(ctx, b, g.L - b) : 0;
                TRY(launch_sweep(ctx, cur, nxt, 0, b, chk_mode, 0, ctx->s_halo, chk_nb, 0));
                TRY(launch_sweep(ctx, cur, nxt, g.L - b, g.L, chk_mode, 0, ctx->s_halo, chk_nb, nb_lo));
                if (fused) {  // rare odd tail step: plain peer copies of the boundary planes
                    const size_t bytes = (size_t)b * g.plane * sizeof(double);
                    if (plo) CK(cudaMemcpyAsync(plo + g.off(ctx->peer_L[0], -g.e, 0), nxt + g.off(0, -g.e, 0), bytes, cudaMemcpyDeviceToDevice, ctx->s_halo));
                    if (phi_) CK(cudaMemcpyAsync(phi_ + g.off(-b, -g.e, 0), nxt + g.off(g.L - b, -g.e, 0), bytes, cudaMemcpyDeviceToDevice, ctx->s_halo));
                }
            }
            if (fused) {
                ctx->pass += 1;
                p2p_signal_kernel<<<1, 1, 0, ctx->s_halo>>>(has_lo ? ctx->peer_flags[0] : nullptr, has_hi ? ctx->peer_flags[1] : nullptr, ctx->pass);
                TRY(post_launch(ctx));
            } else {
                TRY(exchange(ctx, nxt, ctx->s_halo));
            }
            CK(cudaEventRecord(ctx->ev_halo, ctx->s_halo));
            if (two) TRY(launch_sweep_tb2(ctx, src, b, g.L - b, ctx->s_main));
            else TRY(launch_sweep(ctx, cur, nxt, b, g.L - b, chk_mode, 0, ctx->s_main, chk_nb, chk_nb - sweep_blocks(ctx, b, g.L - b)));
            CK(cudaEventRecord(ctx->ev_main, ctx->s_main));
        } else {
            // excited states: sum psi'^2 and (TMA kernel) the overlaps with up to four stored states ride on the sweep
            const int nred = !excited ? chk_mode : (ctx->use_t1 ? 1 + std::min(wnum, t1::MAX_FUSED_LOWERS) : 1);
            if (chk_mode) chk_nb = sweep_blocks(ctx, 0, g.L);
            if (two) TRY(launch_sweep_tb2(ctx, src, 0, g.L, ctx->s_main));
            else TRY(launch_sweep(ctx, cur, nxt, 0, g.L, nred, 0, ctx->s_main));
            TRY(exchange(ctx, nxt, ctx->s_main));
            ctx->cur ^= 1;
            // grid.rs:674-681: norm2 -> normalise -> orthogonalise, every step
            if (excited) TRY(gs_apply(ctx, nxt, ctx->scal + SL_RAW, wnum, nred, sweep_blocks(ctx, 0, g.L)));
        }
        if (overlap) ctx->cur ^= 1;
        done += two ? 2 : 1;
    }
    if (overlap) {
        static const bool skip_final_wait = getenv("WAFER_DEBUG_SKIP_FINAL_WAIT") && atoi(getenv("WAFER_DEBUG_SKIP_FINAL_WAIT")) == 1;
        if (fused && !skip_final_wait) {  // the env knob exists only so that the drift test can prove it catches the race
            // The neighbours' boundary stores of the LAST pass must have landed in my ghost planes before anything
            // that follows on the main stream reads them (observables, Gram-Schmidt, downloads, w_store.push) or
            // overwrites them (normalise, uploads): wait for their pass counter to reach mine.  Without this a rank
            // that runs ahead returns with ghost planes that are one same-buffer pass (4 steps) old.
            p2p_wait_kernel<<<1, 1, 0, ctx->s_halo>>>(ctx->flags, ctx->pass, has_lo, has_hi, ctx->p2p_timeout);
            TRY(post_launch(ctx));
            CK(cudaEventRecord(ctx->ev_halo, ctx->s_halo));
        }
        CK(cudaStreamWaitEvent(ctx->s_main, ctx->ev_halo, 0));
    }
    if (fuse_chk && chk_nb > 0) {
        gs_coeff_kernel<true, false><<<1, 1024, 0, ctx->s_main>>>(ctx->partials, chk_nb, 3, ctx->scal + SL_CHK, 0, nullptr, nullptr,
                                                                    nullptr, 0, nullptr);
        TRY(post_launch(ctx));
        ctx->chk_valid = true;
    }
    return WAFER_OK;
}

int wafer_push_lower(wafer_ctx* ctx, const double* q_padded) {
    try {
        return wafer_push_lower_impl(ctx, q_padded);
    } catch (const std::exception& e) {  // host-side allocation failures must not unwind through the C ABI
        if (ctx) ctx->err = std::string("wafer_push_lower: ") + e.what();
        return WAFER_ERR_INVALID;
    }
}

int wafer_push_lower_from_phi(wafer_ctx* ctx) {
    try {
        return wafer_push_lower_from_phi_impl(ctx);
    } catch (const std::exception& e) {  // host-side allocation failures must not unwind through the C ABI
        if (ctx) ctx->err = std::string("wafer_push_lower_from_phi: ") + e.what();
        return WAFER_ERR_INVALID;
    }
}

int wafer_evolve(wafer_ctx* ctx, uint8_t wnum_in, uint64_t steps) {
    try {
        return wafer_evolve_impl(ctx, wnum_in, steps);
    } catch (const std::exception& e) {  // host-side allocation failures must not unwind through the C ABI
        if (ctx) ctx->err = std::string("wafer_evolve: ") + e.what();
        return WAFER_ERR_INVALID;
    }
}

int wafer_check(wafer_ctx* ctx, uint8_t wnum_in, wafer_observables* out) {
    if (!ctx) return WAFER_ERR_INVALID;
    REQUIRE(out, "out is NULL");
    TRY(ready(ctx, true));
    CK(cudaSetDevice(ctx->dev));
    const int wnum = std::min<int>(wnum_in, (int)ctx->lowers.size());
    TRY(observables_device(ctx));                // grid.rs:127
    TRY(gs_apply(ctx, ctx->psi[ctx->cur], ctx->scal + SL_OBS + 1, wnum));  // grid.rs:130 normalise(norm2), 133-135 orthogonalise
    ctx->chk_valid = false;
    double s[4];
    TRY(read_scalars(ctx, SL_OBS, 4, s));
    out->energy = s[0]; out->norm2 = s[1]; out->v_infinity = s[2]; out->r2 = s[3];
    return WAFER_OK;
}

// grid.rs:50-246
int wafer_solve(wafer_ctx* ctx, uint8_t wnum, double tolerance, int64_t max_steps, uint64_t screen_update,
                uint64_t snap_update, wafer_record* records, uint64_t max_records, uint64_t* n_records) {
    if (!ctx) return WAFER_ERR_INVALID;
    TRY(ready(ctx, true));
    uint64_t step = 0, nrec = 0;
    double last_energy = 1.7976931348623157e308;  // f64::MAX (grid.rs:124)
    bool converged = false;
    for (;;) {
        wafer_observables obs;
        TRY(wafer_check(ctx, wnum, &obs));                                  // grid.rs:127-135
        if (!std::isfinite(obs.energy) || !std::isfinite(obs.norm2) || obs.norm2 == 0.0) {
            ctx->err = "non-finite observables (the reference's R64 would panic here)";
            return WAFER_ERR_NONFINITE;
        }
        const double norm_energy = obs.energy / obs.norm2;                  // grid.rs:128
        const double tau = (double)step * ctx->p.dt;                        // grid.rs:129
        if (snap_update != 0 && step % snap_update == 0)                    // grid.rs:137-139 (NotConstrained):
            TRY(wafer_normalise(ctx, obs.norm2));                           // second division by sqrt(norm2)
        const double diff = std::fabs(norm_energy - last_energy);           // grid.rs:161
        if (records && nrec < max_records) {
            records[nrec].step = step; records[nrec].tau = tau; records[nrec].diff = diff; records[nrec].obs = obs;
        }
        nrec++;
        if (diff < tolerance) { converged = true; break; }                  // grid.rs:162-192
        last_energy = norm_energy;                                          // grid.rs:194
        if (max_steps >= 0 && step > (uint64_t)max_steps) break;            // grid.rs:211-213 (strict >)
        TRY(wafer_evolve(ctx, wnum, screen_update));                        // grid.rs:216
        step += screen_update;                                              // grid.rs:220
    }
    if (n_records) *n_records = nrec;
    if (converged) {
        TRY(wafer_push_lower_from_phi(ctx));                                // grid.rs:241
        return WAFER_OK;
    }
    ctx->err = "Maximum step limit reached before convergence (ErrorKind::MaxStep)";
    return WAFER_ERR_MAX_STEP;
}

// ---- plumbing ----------------------------------------------------------------------------------------------
int wafer_synchronize(wafer_ctx* ctx) {
    if (!ctx) return WAFER_ERR_INVALID;
    CK(cudaSetDevice(ctx->dev));
    CK(cudaStreamSynchronize(ctx->s_halo));
    CK(cudaStreamSynchronize(ctx->s_main));
    return check_p2p_timeout(ctx);
}

int wafer_timer_begin(wafer_ctx* ctx) {
    if (!ctx) return WAFER_ERR_INVALID;
    CK(cudaSetDevice(ctx->dev));
    CK(cudaEventRecord(ctx->ev_t0, ctx->s_main));
    return WAFER_OK;
}

int wafer_timer_end(wafer_ctx* ctx, double* ms) {
    if (!ctx) return WAFER_ERR_INVALID;
    REQUIRE(ms, "ms is NULL");
    CK(cudaSetDevice(ctx->dev));
    CK(cudaEventRecord(ctx->ev_t1, ctx->s_main));
    CK(cudaEventSynchronize(ctx->ev_t1));
    float f = 0.f;
    CK(cudaEventElapsedTime(&f, ctx->ev_t0, ctx->ev_t1));
    *ms = (double)f;
    return WAFER_OK;
}

uint64_t wafer_kernel_launches(const wafer_ctx* ctx) { return ctx ? ctx->launches : 0; }

int wafer_host_alloc(void** ptr, size_t bytes) {
    if (!ptr) return WAFER_ERR_INVALID;
    return cudaMallocHost(ptr, bytes) == cudaSuccess ? WAFER_OK : WAFER_ERR_CUDA;
}

int wafer_host_free(void* ptr) { return cudaFreeHost(ptr) == cudaSuccess ? WAFER_OK : WAFER_ERR_CUDA; }

int wafer_host_register(void* ptr, size_t bytes) {
    if (!ptr || !bytes) return WAFER_ERR_INVALID;
    return cudaHostRegister(ptr, bytes, cudaHostRegisterDefault) == cudaSuccess ? WAFER_OK : WAFER_ERR_CUDA;
}

int wafer_host_unregister(void* ptr) { return cudaHostUnregister(ptr) == cudaSuccess ? WAFER_OK : WAFER_ERR_CUDA; }

int wafer_device_info(const wafer_ctx* ctx, char* name, size_t name_len, int32_t* sm_count, int32_t* cc_major,
                      int32_t* cc_minor, uint64_t* mem_bytes) {
    if (!ctx) return WAFER_ERR_INVALID;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, ctx->dev) != cudaSuccess) return WAFER_ERR_CUDA;
    if (name && name_len) { strncpy(name, prop.name, name_len - 1); name[name_len - 1] = 0; }
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    if (mem_bytes) *mem_bytes = prop.totalGlobalMem;
    return WAFER_OK;
}

// ---- fused halo over peer memory: export / import CUDA IPC handles of the psi buffers and the flag words ------
int wafer_p2p_export(wafer_ctx* ctx, uint8_t out[192]) {
    if (!ctx) return WAFER_ERR_INVALID;
    REQUIRE(out, "out is NULL");
    CK(cudaSetDevice(ctx->dev));
    if (!ctx->flags) {
        CK(cudaMalloc(&ctx->flags, 2 * sizeof(unsigned long long)));
        CK(cudaMemset(ctx->flags, 0, 2 * sizeof(unsigned long long)));
        CK(cudaMalloc(&ctx->p2p_timeout, sizeof(int)));
        CK(cudaMemset(ctx->p2p_timeout, 0, sizeof(int)));
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h[3];
    CK(cudaIpcGetMemHandle(&h[0], ctx->psi[0]));
    CK(cudaIpcGetMemHandle(&h[1], ctx->psi[1]));
    CK(cudaIpcGetMemHandle(&h[2], ctx->flags));
    memcpy(out, h, 192);
    return WAFER_OK;
}

int wafer_p2p_connect(wafer_ctx* ctx, const uint8_t* lower, const uint8_t* upper) {
    if (!ctx) return WAFER_ERR_INVALID;
    REQUIRE(ctx->world > 1, "wafer_p2p_connect needs world > 1");
    REQUIRE(ctx->flags, "call wafer_p2p_export first");
    REQUIRE((ctx->rank == 0) == (lower == nullptr) && (ctx->rank == ctx->world - 1) == (upper == nullptr),
            "pass the export blobs of rank-1 and rank+1 (NULL at the ends of the chain)");
    CK(cudaSetDevice(ctx->dev));
    REQUIRE(!ctx->p2p, "wafer_p2p_connect was already called on this context");
    const uint8_t* blobs[2] = {lower, upper};
    for (int n = 0; n < 2; ++n) {
        if (!blobs[n]) continue;
        cudaIpcMemHandle_t h[3];
        memcpy(h, blobs[n], 192);
        void* p[3];
        for (int k = 0; k < 3; ++k) CK(cudaIpcOpenMemHandle(&p[k], h[k], cudaIpcMemLazyEnablePeerAccess));
        ctx->peer_psi[n][0] = (double*)p[0];
        ctx->peer_psi[n][1] = (double*)p[1];
        ctx->peer_flags[n] = (unsigned long long*)p[2];
        uint64_t x0, x1;
        wafer_slab_partition(ctx->p.nx, (uint32_t)ctx->world, (uint32_t)(ctx->rank + (n == 0 ? -1 : 1)), &x0, &x1);
        ctx->peer_L[n] = (int)(x1 - x0);
    }
    ctx->p2p = true;
    ctx->pass = 0;
    return WAFER_OK;
}

int wafer_selftest_division(wafer_ctx* ctx, double den, uint64_t n, uint64_t seed, uint64_t* mismatches) {
    if (!ctx) return WAFER_ERR_INVALID;
    REQUIRE(mismatches, "mismatches is NULL");
    CK(cudaSetDevice(ctx->dev));
    unsigned long long* d = nullptr;
    CK(cudaMalloc(&d, sizeof(unsigned long long)));
    CK(cudaMemsetAsync(d, 0, sizeof(unsigned long long), ctx->s_main));
    const int ok = (den > 7.888609052210118e-31 && den < 1.2676506002282294e30) ? 1 : 0;
    tb::div_selftest_kernel<<<ctx->sm_count * 8, 256, 0, ctx->s_main>>>(den, ok, n, seed, d);
    TRY(post_launch(ctx));
    unsigned long long h = 0;
    CK(cudaMemcpyAsync(&h, d, sizeof(h), cudaMemcpyDeviceToHost, ctx->s_main));
    CK(cudaStreamSynchronize(ctx->s_main));
    cudaFree(d);
    *mismatches = h;
    return WAFER_OK;
}

int wafer_phi_checksum(wafer_ctx* ctx, uint64_t x_begin, uint64_t x_end, uint64_t out[2]) {
    if (!ctx) return WAFER_ERR_INVALID;
    REQUIRE(out, "out is NULL");
    TRY(ready(ctx, false));
    CK(cudaSetDevice(ctx->dev));
    const Geom& g = ctx->g;
    if (!ctx->d_cksum) CK(cudaMalloc(&ctx->d_cksum, 2 * sizeof(unsigned long long)));
    CK(cudaMemsetAsync(ctx->d_cksum, 0, 2 * sizeof(unsigned long long), ctx->s_main));
    const long long xb = std::max<long long>((long long)x_begin, g.x0) - g.x0;
    const long long xe = std::min<long long>((long long)std::min<uint64_t>(x_end, (uint64_t)g.gnx), g.x0 + g.L) - g.x0;
    if (xe > xb) {
        const int grid = (int)std::min<long long>((xe - xb) * g.ny, (long long)ctx->sm_count * 16);
        checksum_kernel<<<grid, 256, 0, ctx->s_main>>>(ctx->psi[ctx->cur], g, (int)xb, (int)xe, ctx->d_cksum);
        TRY(post_launch(ctx));
    }
    unsigned long long h[2] = {0, 0};
    CK(cudaMemcpyAsync(h, ctx->d_cksum, sizeof(h), cudaMemcpyDeviceToHost, ctx->s_main));
    CK(cudaStreamSynchronize(ctx->s_main));
    out[0] = h[0]; out[1] = h[1];
    return check_p2p_timeout(ctx);
}

int wafer_debug_halo_delay(wafer_ctx* ctx, uint64_t nanoseconds) {
    if (!ctx) return WAFER_ERR_INVALID;
    REQUIRE(nanoseconds <= 1000000000ull, "delay is capped at one second per pass");
    ctx->dbg_halo_delay_ns = nanoseconds;
    return WAFER_OK;
}

const char* wafer_sweep_variant(const wafer_ctx* ctx) {
    if (!ctx) return "";
    if (ctx->use_tb) return ctx->p2p ? "tb2-tma/V-onfly+p2p-halo" : (ctx->use_t1 ? "tb2-tma/V-onfly+tma1" : "tb2-tma/V-onfly");
    if (ctx->use_t1) return "tma1/V-onfly";
    return ctx->onfly ? "simple-regqueue/V-onfly" : "simple-regqueue/AB-arrays";
}

}  // extern "C"
