// Host side of libmia_scan.so: argument validation (the reference's TORCH_CHECKs,
// selective_scan_oflex.cpp:152-204 / 245-312), tile planning, launches, the finalize kernel that folds the
// backward's deterministic partials, and the extern "C" surface declared in include/mia_selective_scan.h.
#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../include/mia_selective_scan.h"
#include "scan_common.cuh"
#include "scan_bwd_rows.cuh"
#include "scan_bwd_rowsn.cuh"
#include "scan_bwd_cw.cuh"
#include "tma_host.h"
#include "scan_fwd_rowsn.cuh"
#include "scan_fwd_stream.cuh"
#include "scan_fwd_chunks.cuh"
#include "scan_fwd_cw.cuh"

namespace mia {
template <typename T> cudaError_t launch_fwd_rows(const RowsArgs &, int, bool, cudaStream_t);
template <typename T> cudaError_t launch_fwd_rowsn(const RowsNArgs &, int, bool, cudaStream_t);
template <typename T> cudaError_t launch_fwd_stream(const StreamArgs &, int, bool, cudaStream_t);
template <typename T> cudaError_t launch_fwd_chunks(const ChunkArgs &, int, bool, cudaStream_t);
template <typename T> cudaError_t launch_fwd_cw(const CUtensorMap *, const CwFwdArgs &, int, bool, cudaStream_t);
template <typename T> cudaError_t launch_bwd_rows(const RowsBwdArgs &, int, bool, cudaStream_t);
template <typename T> cudaError_t launch_bwd_rowsn(const RowsNBwdArgs &, int, bool, cudaStream_t);
template <typename T> cudaError_t launch_bwd_cw(const CUtensorMap *, const CwBwdArgs &, int, bool, cudaStream_t);
template <typename T> cudaError_t launch_fwd_any(const ScanArgs &, int, cudaStream_t);
template <typename T> cudaError_t launch_bwd_any(const ScanArgs &, int, cudaStream_t);
}  // namespace mia

namespace {

thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};

int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define MIA_CHECK(cond, ...) \
    do { if (!(cond)) return fail(MIA_EINVAL, __VA_ARGS__); } while (0)
#define MIA_CUDA(expr) \
    do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) return fail(MIA_ECUDA, "%s: %s", #expr, cudaGetErrorString(e__)); } while (0)

// Debugging knobs (force a kernel family, move the half split) exist only in -DMIA_DEBUG builds; the release library
// reads no environment variable and has no mode that alters results.
#ifdef MIA_DEBUG
bool dbg_knob(const char *name) { return getenv(name) != nullptr; }
int dbg_int(const char *name, int dflt) { const char *v = getenv(name); return v ? atoi(v) : dflt; }
#define MIA_TRACE(...) do { if (getenv("MIA_TRACE")) { fprintf(stderr, "[mia] " __VA_ARGS__); fputc('\n', stderr); } } while (0)
#else
#define MIA_TRACE(...) do { } while (0)
constexpr bool dbg_knob(const char *) { return false; }
constexpr int dbg_int(const char *, int dflt) { return dflt; }
#endif

int esize(int dt) { return dt == MIA_F32 ? 4 : 2; }
int round_up(int x, int m) { return (x + m - 1) / m * m; }

struct DeviceInfo { int sms = 0; int smem_optin = 0; };
int device_info(DeviceInfo &di) {
    static thread_local int cached_dev = -1;
    static thread_local DeviceInfo cached;
    int dev = 0;
    MIA_CUDA(cudaGetDevice(&dev));
    if (dev != cached_dev) {
        MIA_CUDA(cudaDeviceGetAttribute(&cached.sms, cudaDevAttrMultiProcessorCount, dev));
        MIA_CUDA(cudaDeviceGetAttribute(&cached.smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
        cached_dev = dev;
    }
    di = cached;
    return MIA_OK;
}

int lanes_per_row(int L) {
    if (L > 128) return 32;
    int need = (L + mia::kTok - 1) / mia::kTok, lpr = 4;
    while (lpr < need) lpr <<= 1;
    return lpr;
}

struct WorkspaceLayout {
    size_t part_dA = 0, part_dD = 0, part_dbias = 0, acc_dB = 0, acc_dC = 0, ddelta_full = 0, total = 0;
    size_t acc_bytes = 0;
    bool bc_atomic = false;
};

// Work plan shared by fwd and bwd (the checkpoint geometry must match between them).
struct Plan {
    int LPR, CH, n_chunks, RPP, NW, RT, RS, split, tiles_per_seg, n_seg;
};

void set_split(Plan &pl, const mia_ss_params &p, int split) {
    const int rpg = p.dim / p.n_groups;
    pl.RS = (rpg + split - 1) / split;
    pl.split = (rpg + pl.RS - 1) / pl.RS;
    pl.tiles_per_seg = (pl.RS + pl.RT - 1) / pl.RT;
    pl.n_seg = p.batch * p.n_groups * pl.split;
}

Plan make_plan(const mia_ss_params &p, int sms, bool bwd) {
    Plan pl;
    pl.LPR = lanes_per_row(p.seqlen);
    pl.CH = pl.LPR * mia::kTok;
    pl.n_chunks = (p.seqlen + pl.CH - 1) / pl.CH;
    pl.RPP = 32 / pl.LPR;
    pl.NW = (bwd ? mia::kThreads : mia::kThreadsFwd) / 32 - 1;
    pl.RT = pl.NW * pl.RPP * (bwd ? 4 : 3);   // row passes per warp and row stage (amortises the per-stage bookkeeping)
    const int rpg = p.dim / p.n_groups;
    // Segments: (batch, group, row range).  Aim at >= ~6 segments per SM with the best last-wave balance; a segment
    // keeps its carries in shared memory, so bound rows * d_state.
    const long long bg = (long long)p.batch * p.n_groups;
    int max_split = (rpg + pl.RPP - 1) / pl.RPP;
    int min_split = 1;
    while ((long long)((rpg + min_split - 1) / min_split) * p.dstate > 4096 && min_split < max_split) ++min_split;
    int want = (int)((6LL * sms + bg - 1) / bg);
    if (want < min_split) want = min_split;
    if (want > max_split) want = max_split;
    int best = want;
    double best_eff = -1.0;
    for (int sp = want; sp <= want + 3 && sp <= max_split; ++sp) {
        Plan t = pl;
        set_split(t, p, sp);
        const double waves = (double)t.n_seg / sms;
        const double eff = waves / (double)((t.n_seg + sms - 1) / sms);
        if (eff > best_eff + 1e-9) { best_eff = eff; best = sp; }
    }
    set_split(pl, p, best);
    return pl;
}

// Work split of the deterministic d_state 16 backward (scan_bwd_rowsn.cuh): sizes only, shared by the workspace query and
// the launch.  Returns false when the shape is not one it takes.
struct RowsNSplit { int n_oct, oct_per_cta, oct_per_group, grid, max_parts; };
bool rowsn_bwd_split(const mia_ss_params &p, int sms, int ctas_per_sm, RowsNSplit &sp) {
    const int rpg = p.dim / p.n_groups;
    if (p.dstate != 16 || p.delta_dim != p.dim || p.seqlen > 256 || (rpg % mia::kRnOct)) return false;
    sp.n_oct = (int)((long long)p.batch * p.dim / mia::kRnOct);
    sp.oct_per_group = rpg / mia::kRnOct;
    const int slots = sms * ctas_per_sm;
    sp.oct_per_cta = (sp.n_oct + slots - 1) / slots;
    sp.grid = (sp.n_oct + sp.oct_per_cta - 1) / sp.oct_per_cta;
    sp.max_parts = (sp.oct_per_group + sp.oct_per_cta - 2) / sp.oct_per_cta + 1;   // CTAs a group's octets can straddle
    return true;
}
constexpr int kRowsNCtasPerSm = 3;   // the split (hence the partial layout) is fixed at 3 CTAs per SM: launching fewer
                                     // resident CTAs (z gate, fp32) only queues them

WorkspaceLayout workspace_layout(const mia_ss_params &p, const Plan &pl, int sms) {
    WorkspaceLayout w;
    auto take = [&](size_t nfloat) { size_t off = w.total; w.total += (nfloat * 4 + 255) / 256 * 256; return off; };
    w.bc_atomic = p.dstate > 1;
    w.part_dA = take((size_t)p.batch * p.dim * p.dstate);
    w.part_dD = take((size_t)p.batch * p.dim);
    w.part_dbias = take((size_t)p.batch * p.dim);
    // d_state == 1: one partial row per segment (warp-scan kernels) or per 32-row batch (row-serial kernel); the query
    // cannot know which of the two will run (it depends on strides and alignment), so size for the larger.
    const int rpg = p.dim / p.n_groups;
    const size_t row_items = (p.dstate == 1 && rpg % 32 == 0) ? (size_t)p.batch * p.n_groups * (rpg / 32) : 0;
    const size_t parts = row_items > (size_t)pl.n_seg ? row_items : (size_t)pl.n_seg;
    size_t nacc = w.bc_atomic ? (size_t)p.batch * p.n_groups * p.dstate * ((p.seqlen + 3) & ~3) : parts * p.dstate * p.seqlen;
    RowsNSplit rs;
    if (rowsn_bwd_split(p, sms, kRowsNCtasPerSm, rs)) {   // whichever d_state 16 kernel runs (it depends on strides), its buffer fits
        const size_t need = (size_t)p.batch * p.n_groups * rs.max_parts * p.dstate * p.seqlen;
        if (need > nacc) nacc = need;
    }
    w.acc_dB = take(nacc);
    w.acc_dC = take(nacc);
    w.acc_bytes = nacc * 4;
    if (p.dim != p.delta_dim) w.ddelta_full = take((size_t)p.batch * p.dim * p.seqlen);
    return w;
}

int validate_sizes(const mia_ss_params &p) {
    MIA_CHECK(p.itype == MIA_F32 || p.itype == MIA_F16 || p.itype == MIA_BF16, "u must be float32, float16 or bfloat16");
    MIA_CHECK(p.otype == p.itype || p.otype == MIA_F32, "out/dout dtype must be the input dtype or float32");
    MIA_CHECK(p.batch > 0 && p.dim > 0 && p.seqlen > 0 && p.dstate > 0 && p.n_groups > 0 && p.delta_dim > 0, "empty or negative size");
    MIA_CHECK(p.dim % p.n_groups == 0, "dims should be dividable by n_groups");
    MIA_CHECK(p.dim % p.delta_dim == 0, "dims should be dividable by delta_dim");
    MIA_CHECK(p.dstate <= 256, "selective_scan only supports state dimension <= 256");
    MIA_CHECK((long long)p.batch * p.dim < (1LL << 31) / 2, "batch*dim too large");
    return MIA_OK;
}

int validate_common(const mia_ss_params &p) {
    if (int rc = validate_sizes(p)) return rc;
    MIA_CHECK(p.u && p.delta && p.A && p.B && p.C, "u, delta, A, B, C must not be null");
    MIA_CHECK(p.n_chunks == mia_ss_num_chunks(p.seqlen), "x must have %d chunks for seqlen %d (got %d)", mia_ss_num_chunks(p.seqlen),
              p.seqlen, p.n_chunks);
    return MIA_OK;
}

void fill_common(mia::ScanArgs &a, const mia_ss_params &p, const Plan &pl, bool bwd) {
    memset(&a, 0, sizeof(a));
    a.batch = p.batch; a.dim = p.dim; a.L = p.seqlen; a.N = p.dstate; a.G = p.n_groups; a.delta_dim = p.delta_dim;
    a.rows_per_group = p.dim / p.n_groups;
    a.delta_ratio = p.dim / p.delta_dim;
    a.softplus = p.delta_softplus; a.has_z = p.z != nullptr; a.out_f32 = (p.otype == MIA_F32) && (p.itype != MIA_F32);
    a.RT = pl.RT; a.RS = pl.RS; a.split = pl.split; a.tiles_per_seg = pl.tiles_per_seg; a.LPR = pl.LPR; a.CH = pl.CH;
    a.n_chunks = pl.n_chunks; a.n_seg = pl.n_seg; a.n_consumer_warps = pl.NW;
    a.u = p.u; a.delta = p.delta; a.A = p.A; a.B = p.B; a.C = p.C; a.D = p.D; a.delta_bias = p.delta_bias; a.z = p.z;
    a.x = p.x;
    a.u_bs = p.u_batch_stride; a.u_ds = p.u_d_stride; a.delta_bs = p.delta_batch_stride; a.delta_ds = p.delta_d_stride;
    a.z_bs = p.z_batch_stride; a.z_ds = p.z_d_stride; a.A_ds = p.A_d_stride; a.A_ns = p.A_dstate_stride;
    a.B_bs = p.B_batch_stride; a.B_gs = p.B_group_stride; a.B_ns = p.B_dstate_stride;
    a.C_bs = p.C_batch_stride; a.C_gs = p.C_group_stride; a.C_ns = p.C_dstate_stride;
    const bool whole = pl.n_chunks == 1;
    a.flat_u = whole && p.u_d_stride == p.seqlen;
    a.flat_delta = whole && p.delta_d_stride == p.seqlen;
    a.flat_z = a.has_z && whole && p.z_d_stride == p.seqlen;
    a.flat_B = whole && p.B_dstate_stride == p.seqlen;
    a.flat_C = whole && p.C_dstate_stride == p.seqlen;
    (void)bwd;
}

// Shared-memory carve-up.  Returns false if not even two row stages fit.
bool layout_smem(mia::ScanArgs &a, int es, int eso, bool bwd, int smem_max) {
    const int span = a.CH < a.L ? a.CH : a.L;
    a.row_pitch = round_up(span * es, 16) + 16;
    a.rowo_pitch = round_up(span * eso, 16) + 16;
    a.bc_pitch = a.row_pitch;
    const int slack = a.CH * 4 + 16;   // an 8-token vector read may run past the last row of a region
    int off = 0;
    auto region = [&](int rows, int pitch) { int o = off; off += round_up(rows * pitch + slack, 128); return o; };
    a.off_u = region(a.RT, a.row_pitch);
    a.off_delta = region(a.RT, a.row_pitch);
    if (a.has_z) a.off_z = region(a.RT, a.row_pitch);
    if (bwd) {
        a.off_dout = region(a.RT, a.rowo_pitch);
        if (a.has_z) a.off_osaved = region(a.RT, a.rowo_pitch);
        a.off_h0 = off; off += round_up(a.RT * a.N * 4, 128);
    }
    a.stage_bytes = off;
    off = 0;
    a.goff_B = region(a.N, a.bc_pitch);
    a.goff_C = region(a.N, a.bc_pitch);
    a.goff_A = off; off += round_up(a.RS * a.N * 4, 128);
    a.goff_D = off; off += round_up(a.RS * 4, 128);
    a.goff_bias = off; off += round_up(a.RS * 4, 128);
    a.gstage_bytes = off;
    const int bars = round_up((2 * mia::kMaxStages + 2 * mia::kGroupStages) * 8, 128);
    const int carry = round_up(bwd ? (2 * a.RS * a.N + 2 * a.RS) * 4 : a.RS * a.N * 8, 128);
    const int red = bwd ? a.n_consumer_warps * 256 * 4 : 0;
    const int bcf = 0;   // (an fp32 copy of the B / C chunk for the d_state > 1 fast backward was tried and removed)
    const int fixed = mia::kGroupStages * a.gstage_bytes + bars + carry + red + bcf;
    int stages = (smem_max - fixed) / a.stage_bytes;
    if (stages > mia::kMaxStages) stages = mia::kMaxStages;
    if (stages < 2) return false;
    a.stages = stages;
    a.off_groups = stages * a.stage_bytes;
    a.off_bars = a.off_groups + mia::kGroupStages * a.gstage_bytes;
    a.off_carry = a.off_bars + bars;
    a.off_red = a.off_carry + carry;
    a.off_bcf = bcf ? a.off_red + red : 0;
    a.smem_bytes = a.off_red + red + bcf;
    return true;
}

// Pick the work plan and carve shared memory; if it does not fit, first cut segments finer (smaller carries and
// parameter blocks), then halve the row tile.
int plan_and_layout(const mia_ss_params &p, const DeviceInfo &di, bool bwd, Plan &pl, mia::ScanArgs &a) {
    pl = make_plan(p, di.sms, bwd);
    const int es = esize(p.itype), eso = esize(p.otype);
    const int rpg = p.dim / p.n_groups;
    for (;;) {
        fill_common(a, p, pl, bwd);
        if (bwd) {
            const bool whole = pl.n_chunks == 1;
            a.flat_dout = whole && p.dout_d_stride == p.seqlen;
            a.flat_osaved = a.has_z && whole && p.out_saved_d_stride == p.seqlen;
        }
        if (layout_smem(a, es, eso, bwd, di.smem_optin)) return MIA_OK;
        if (pl.RS > pl.RT) {
            set_split(pl, p, pl.split * 2 < rpg ? pl.split * 2 : rpg);
        } else {
            MIA_CHECK(pl.RT > pl.RPP, "tile does not fit in shared memory (dstate %d, seqlen %d)", p.dstate, p.seqlen);
            pl.RT = pl.RT / 2 < pl.RPP ? pl.RPP : pl.RT / 2;
            set_split(pl, p, pl.split);
        }
    }
}

// Geometry of the column-walk kernels (scan_fwd_cw.cuh / scan_bwd_cw.cuh): g rows per tensor-map row.  A pure function of the
// problem's sizes and dtypes (and the SM count) -- it also fixes the LAYOUT of the block states (hblk), which forward and backward must agree on.
// A tensor-map row must be a multiple of 16 bytes (else the map cannot be encoded); it SHOULD be a multiple of 32 (else every
// other row starts mid-sector and every 64-byte box row touches three sectors instead of two: +19 % DRAM reads measured at
// L = 196 bf16 with g = 2).  2-byte types: the smallest g in {1, 2, 4} with g L a multiple of 16 elements, falling back to a
// smaller legal g when rows_per_group is not a multiple of 32 g or when the item count falls between one and two rounds; g > 1
// only for rows of one checkpoint chunk (L <= 256: the kernels write x at row ends only when they walk several rows per
// lane).  fp32 rows (L % 4 == 0) are always legal: g = 1.
bool cw_geometry(const mia_ss_params &p, int sms, int &g) {
    const int es = esize(p.itype), L = p.seqlen;
    if (p.dstate != 1 || p.z || p.delta_dim != p.dim || (L % 4) || p.n_groups < 1) return false;
    const int rpg = p.dim / p.n_groups;
    if (es == 4) { g = 1; return rpg % 32 == 0; }
    const int want = (L % 16 == 0) ? 1 : (L % 8 == 0) ? 2 : 4;
    // Items are 32 g rows: a larger g halves their number, and a problem that fills the resident slots a little more than once
    // pays a whole second round (measured, gpurun r2fin2, B = 64, L = 196, fp32 out: 1536 items on 1332 - 1480 slots, forward
    // 0.079 -> 0.103 ms) -> take the preferred g when its items fit one round or make at least two, else the next smaller
    // legal one.  (~9 resident warps per SM with fp32 outputs / gradients, 12 otherwise; the same estimate for both directions,
    // since forward and backward must pick the same g.)
    const long long slots = (long long)sms * ((esize(p.otype) == 4) ? 9 : 12);
    int fallback = 0;
    for (int c = want; c >= 1; c >>= 1) {
        if (((long long)c * L * es) % 16) break;                 // not a legal tensor-map row any more
        if (c > 1 && L > 256) continue;
        if (rpg % (32 * c)) continue;
        const long long n_items = (long long)p.batch * p.n_groups * (rpg / (32 * c));
        if (n_items <= slots || n_items >= 2 * slots) { g = c; return true; }
        if (!fallback) fallback = c;
    }
    if (fallback) { g = fallback; return true; }
    return false;
}

// Row-serial forward (scan_fwd_rows.cuh): eligibility + argument block.  Returns false when the warp-scan kernels must run.
bool plan_rows_fwd(const mia_ss_params &p, const DeviceInfo &di, mia::RowsArgs &r, int &grid) {
    const int es = esize(p.itype), L = p.seqlen;
    const int rpg = p.dim / p.n_groups;
    if (p.dstate != 1 || p.z || p.delta_dim != p.dim || (rpg % 32) || (L % 4)) return false;
    if (dbg_knob("MIA_NO_ROWS_FWD")) return false;
    auto dense = [&](long long bs, long long ds) { return ds == L && bs == (long long)p.dim * L; };
    if (!dense(p.u_batch_stride, p.u_d_stride) || !dense(p.delta_batch_stride, p.delta_d_stride) ||
        !dense(p.out_batch_stride, p.out_d_stride)) return false;
    if (p.A_d_stride != 1 && p.dim > 1) return false;
    if (((uintptr_t)p.u | (uintptr_t)p.delta | (uintptr_t)p.out) & 15) return false;
    memset(&r, 0, sizeof(r));
    const int budget = 13 * 1024;                               // bytes of one [32 x chunk] tile (8 warps per SM at 2-byte dtypes)
    int Lc = L;
    if (32 * L * es > budget) {
        if ((L * es) % 16) return false;                        // per-row pieces must start 16-byte aligned
        Lc = budget / (32 * es) / 8 * 8;
    }
    r.batch = p.batch; r.dim = p.dim; r.L = L; r.G = p.n_groups; r.rows_per_group = rpg; r.softplus = p.delta_softplus;
    r.Lc = Lc; r.n_lchunks = (L + Lc - 1) / Lc;
    r.n_items = p.batch * p.n_groups * (rpg / 32);
    r.tile_bytes = round_up(32 * Lc * es, 128);
    r.bc_bytes = 0;
    r.off_bc32 = 2 * r.tile_bytes;
    r.stage_bytes = r.off_bc32 + round_up(2 * Lc * 4, 128);
    r.smem_bytes = r.stage_bytes + 128;
    if (r.smem_bytes > di.smem_optin) return false;
    r.xchunks = mia_ss_num_chunks(L); r.xchunk_tokens = mia_ss_chunk_len(L);
    r.u = p.u; r.delta = p.delta; r.A = p.A; r.B = p.B; r.C = p.C; r.D = p.D; r.delta_bias = p.delta_bias; r.out = p.out; r.x = p.x;
    r.B_bs = p.B_batch_stride; r.B_gs = p.B_group_stride; r.C_bs = p.C_batch_stride; r.C_gs = p.C_group_stride;
    const int per_sm = di.smem_optin / (r.smem_bytes + 1024) > 0 ? (227 * 1024) / (r.smem_bytes + 1024) : 1;
    grid = di.sms * (per_sm < 1 ? 1 : (per_sm > 8 ? 8 : per_sm));
    if (grid > r.n_items) grid = r.n_items;
    return true;
}

// Column-walk forward (scan_fwd_cw.cuh): eligibility, argument block, tensor maps (u, delta, out).
bool plan_cw_fwd(const mia_ss_params &p, const DeviceInfo &di, mia::CwFwdArgs &r, CUtensorMap *tm, int &grid) {
    const int es = esize(p.itype), eo = esize(p.otype), L = p.seqlen;
    int g = 0;
    if (!cw_geometry(p, di.sms, g)) return false;
    if (dbg_knob("MIA_NO_CW_FWD")) return false;
    const int rpg = p.dim / p.n_groups;
    auto dense = [&](long long bs, long long ds) { return ds == L && bs == (long long)p.dim * L; };
    if (!dense(p.u_batch_stride, p.u_d_stride) || !dense(p.delta_batch_stride, p.delta_d_stride) ||
        !dense(p.out_batch_stride, p.out_d_stride)) return false;
    if (p.A_d_stride != 1 && p.dim > 1) return false;
    if (((uintptr_t)p.u | (uintptr_t)p.delta | (uintptr_t)p.out) & 15) return false;
    const unsigned long long trows = (unsigned long long)p.batch * p.dim / g, tcols = (unsigned long long)g * L;
    if (tcols >= (1ull << 31) || trows >= (1ull << 31)) return false;
    const int n_items = p.batch * p.n_groups * (rpg / (32 * g));
    // few long rows: the chunk-parallel kernel (scan_fwd_chunks.cuh) has more warps to offer than one per 32 rows
    if (n_items < 4 * di.sms && (L + mia::kChunkTok - 1) / mia::kChunkTok >= 2 && !dbg_knob("MIA_FORCE_CW_FWD")) return false;
    memset(&r, 0, sizeof(r));
    r.zero = 0;                                                  // (the kernel XORs parameter loads with it: scan_fwd_cw.cuh, settle)
    r.batch = p.batch; r.dim = p.dim; r.L = L; r.G = p.n_groups; r.rows_per_group = rpg; r.softplus = p.delta_softplus;
    r.g = g; r.n_items = n_items;
    r.nwin = (g * L + mia::kCwTok - 1) / mia::kCwTok;
    r.ngrp = (g * L + 15) / 16;
    const bool of32 = eo == 4 && es != 4;
    const int tile_i = 32 * mia::kCwTok * es, tile_o = 32 * mia::kCwTok * eo;
    r.stage_bytes = 2 * tile_i;
    // stages: 4 unless 3 gives a better-balanced last round (more resident warps)
    int best_ns = 0, best_per_sm = 0;
    double best_eff = -1.0;
    for (int ns = 4; ns >= 3; --ns) {
        const int smem = ns * r.stage_bytes + (of32 ? 2 * tile_o : 0) + 2 * mia::kCwTok * 4 + 8 * ns + 16;
        int per_sm = (227 * 1024) / (smem + 1024);
        if (per_sm > 16) per_sm = 16;
        if (per_sm < 1) continue;
        const long long slots = (long long)di.sms * per_sm;
        const long long rounds = (n_items + slots - 1) / slots;
        const double eff = (double)n_items / (double)(rounds * slots) * (per_sm >= 8 ? 1.0 : per_sm / 8.0);
        if (eff > best_eff + 0.03) { best_eff = eff; best_ns = ns; best_per_sm = per_sm; }
    }
    if (!best_ns) return false;
    r.ns = dbg_int("MIA_CW_STAGES", best_ns);
    r.off_out = r.ns * r.stage_bytes;
    r.off_bc32 = r.off_out + (of32 ? 2 * tile_o : 0);
    r.off_bar = r.off_bc32 + 2 * mia::kCwTok * 4;
    r.smem_bytes = r.off_bar + 8 * r.ns + 16;
    r.xchunks = mia_ss_num_chunks(L); r.xchunk_tokens = mia_ss_chunk_len(L);
    r.A = p.A; r.B = p.B; r.C = p.C; r.D = p.D; r.delta_bias = p.delta_bias; r.x = p.x; r.hblk = p.hblk;
    r.B_bs = p.B_batch_stride; r.B_gs = p.B_group_stride; r.C_bs = p.C_batch_stride; r.C_gs = p.C_group_stride;
    const void *ptrs[3] = {p.u, p.delta, p.out};
    for (int i = 0; i < 3; ++i) {
        const int e = i == 2 ? eo : es;
        const int trc = mia::tma_make_2d(&tm[i], ptrs[i], trows, tcols, tcols * e, 32, mia::kCwTok, e, mia::kCwTok * e);
        if (trc != 0) { MIA_TRACE("cw fwd: tensor map %d rejected (CUresult %d)", i, trc); return false; }
    }
    int per_sm = (227 * 1024) / (r.smem_bytes + 1024);
    if (per_sm > 16) per_sm = 16;
    per_sm = std::min(per_sm, dbg_int("MIA_CW_MAXPERSM", per_sm));
    if (per_sm < 1) return false;
    const long long slots = (long long)di.sms * per_sm;
    const long long rounds = (n_items + slots - 1) / slots;
    grid = (int)((n_items + rounds - 1) / rounds);              // equal rounds per CTA
    (void)best_per_sm;
    return true;
}

// Chunk-parallel forward for long rows and few 32-row batches (scan_fwd_chunks.cuh): eligibility + argument block.
bool plan_chunks_fwd(const mia_ss_params &p, const DeviceInfo &di, mia::ChunkArgs &r, int &grid) {
    const int es = esize(p.itype), eo = esize(p.otype), L = p.seqlen;
    const int rpg = p.dim / p.n_groups;
    if (p.dstate != 1 || p.z || p.delta_dim != p.dim || (rpg % 32) || (L % 4)) return false;
    if (dbg_knob("MIA_NO_ROWS_FWD") || dbg_knob("MIA_NO_CHUNKS_FWD")) return false;
    const int nch = (L + mia::kChunkTok - 1) / mia::kChunkTok;
    if (nch < 2 || mia_ss_chunk_len(L) != mia::kChunkTok || ((L * es) % 16) || ((L * eo) % 16)) return false;
    const int n_batches = p.batch * p.n_groups * (rpg / 32);
    if (n_batches >= 4 * di.sms) return false;                 // enough rows: the serial row walk fills the machine
    auto dense = [&](long long bs, long long ds) { return ds == L && bs == (long long)p.dim * L; };
    if (!dense(p.u_batch_stride, p.u_d_stride) || !dense(p.delta_batch_stride, p.delta_d_stride) ||
        !dense(p.out_batch_stride, p.out_d_stride)) return false;
    if (p.A_d_stride != 1 && p.dim > 1) return false;
    if (((uintptr_t)p.u | (uintptr_t)p.delta | (uintptr_t)p.out) & 15) return false;
    memset(&r, 0, sizeof(r));
    r.batch = p.batch; r.dim = p.dim; r.L = L; r.G = p.n_groups; r.rows_per_group = rpg; r.softplus = p.delta_softplus;
    r.n_chunks = nch;
    r.n_items = n_batches * nch;
    r.tile_bytes = round_up(32 * (mia::kChunkTok * es + 16), 128);
    r.off_bc32 = 2 * r.tile_bytes;
    r.off_bar = r.off_bc32 + 2 * mia::kChunkTok * 4;
    r.smem_bytes = r.off_bar + 128;
    int per_sm = (227 * 1024) / (r.smem_bytes + 1024);
    if (per_sm < 3) return false;
    if (per_sm > 8) per_sm = 8;
    r.u = p.u; r.delta = p.delta; r.A = p.A; r.B = p.B; r.C = p.C; r.D = p.D; r.delta_bias = p.delta_bias; r.out = p.out; r.x = p.x;
    r.B_bs = p.B_batch_stride; r.B_gs = p.B_group_stride; r.C_bs = p.C_batch_stride; r.C_gs = p.C_group_stride;
    grid = di.sms * per_sm;
    if (grid > r.n_items) grid = r.n_items;
    return true;
}

// Streaming row-serial forward (scan_fwd_stream.cuh): eligibility + argument block.
bool plan_stream_fwd(const mia_ss_params &p, const DeviceInfo &di, mia::StreamArgs &r, int &grid) {
    const int es = esize(p.itype), eo = esize(p.otype), L = p.seqlen;
    const int rpg = p.dim / p.n_groups;
    if (p.dstate != 1 || p.z || p.delta_dim != p.dim || (rpg % 32) || (L % 4)) return false;
    if (dbg_knob("MIA_NO_ROWS_FWD") || dbg_knob("MIA_NO_STREAM_FWD")) return false;
    auto dense = [&](long long bs, long long ds) { return ds == L && bs == (long long)p.dim * L; };
    if (!dense(p.u_batch_stride, p.u_d_stride) || !dense(p.delta_batch_stride, p.delta_d_stride) ||
        !dense(p.out_batch_stride, p.out_d_stride)) return false;
    if (p.A_d_stride != 1 && p.dim > 1) return false;
    const uintptr_t piece = 4 * es - 1, pieceo = 4 * eo - 1;    // every row start is a multiple of one 4-token piece
    if ((((uintptr_t)p.u | (uintptr_t)p.delta) & piece) || ((uintptr_t)p.out & pieceo)) return false;
    if (L > 256 && mia_ss_chunk_len(L) != 256) return false;
    memset(&r, 0, sizeof(r));
    r.batch = p.batch; r.dim = p.dim; r.L = L; r.G = p.n_groups; r.rows_per_group = rpg; r.softplus = p.delta_softplus;
    r.n_items = p.batch * p.n_groups * (rpg / 32);
    r.Lp = (L + 3) & ~3;
    const int pitch = mia::kStreamTok * es + 4 * es;
    r.off_bc32 = round_up(2 * 2 * 32 * pitch, 128);             // 2 stages x (u, delta)
    r.smem_bytes = r.off_bc32 + round_up(2 * r.Lp * 4, 128);
    if (r.smem_bytes > di.smem_optin) return false;
    r.xchunks = mia_ss_num_chunks(L);
    r.u = p.u; r.delta = p.delta; r.A = p.A; r.B = p.B; r.C = p.C; r.D = p.D; r.delta_bias = p.delta_bias; r.out = p.out; r.x = p.x;
    r.B_bs = p.B_batch_stride; r.B_gs = p.B_group_stride; r.C_bs = p.C_batch_stride; r.C_gs = p.C_group_stride;
    int per_sm = (227 * 1024) / (r.smem_bytes + 1024);
    if (per_sm > 16) per_sm = 16;                               // 128 registers per thread
    if (per_sm < 4) return false;
    grid = di.sms * per_sm;
    if (grid > r.n_items) grid = r.n_items;
    return true;
}

// Row-serial forward for d_state > 1 (scan_fwd_rowsn.cuh): eligibility + argument block.
bool plan_rowsn_fwd(const mia_ss_params &p, const DeviceInfo &di, mia::RowsNArgs &r, int &grid) {
    const int es = esize(p.itype), L = p.seqlen, N = p.dstate;
    const int rpg = p.dim / p.n_groups;
    if ((N != 16 && N != 8) || p.delta_dim != p.dim || (rpg % 32)) return false;
    if (dbg_knob("MIA_NO_ROWS_FWD")) return false;
    if (mia_ss_num_chunks(L) != 1) return false;                // whole rows, one checkpoint
    auto dense = [&](long long bs, long long ds) { return ds == L && bs == (long long)p.dim * L; };
    if (!dense(p.u_batch_stride, p.u_d_stride) || !dense(p.delta_batch_stride, p.delta_d_stride) ||
        !dense(p.out_batch_stride, p.out_d_stride)) return false;
    if (((uintptr_t)p.u | (uintptr_t)p.delta | (uintptr_t)p.out) & 15) return false;
    if (p.z && (!dense(p.z_batch_stride, p.z_d_stride) || !dense(p.out_z_batch_stride, p.out_z_d_stride) ||
                (((uintptr_t)p.z | (uintptr_t)p.out_z) & 15))) return false;
    if ((32 * L * es) % 16) return false;
    memset(&r, 0, sizeof(r));
    r.batch = p.batch; r.dim = p.dim; r.L = L; r.G = p.n_groups; r.rows_per_group = rpg; r.N = N; r.softplus = p.delta_softplus;
    r.warps = p.z ? 2 : mia::kRowsNWarps;
    r.tiles_per_warp = p.z ? 3 : 2;
    const int W = r.warps;
    r.units_per_group = (rpg / 32 + W - 1) / W;
    r.n_units = p.batch * p.n_groups * r.units_per_group;
    r.tile_bytes = round_up(32 * L * es, 128);
    r.off_bc = W * r.tiles_per_warp * r.tile_bytes;
    r.off_bar = r.off_bc + round_up(2 * L * N * es, 128);
    r.smem_bytes = r.off_bar + 128;
    const int per_sm = (int)(233472 / (r.smem_bytes + 1024));
    if (per_sm < 1 || r.smem_bytes > di.smem_optin) return false;
    r.u = p.u; r.delta = p.delta; r.A = p.A; r.B = p.B; r.C = p.C; r.D = p.D; r.delta_bias = p.delta_bias; r.out = p.out; r.x = p.x;
    r.z = p.z; r.out_z = p.out_z;
    r.A_ds = p.A_d_stride; r.A_ns = p.A_dstate_stride;
    r.B_bs = p.B_batch_stride; r.B_gs = p.B_group_stride; r.B_ns = p.B_dstate_stride;
    r.C_bs = p.C_batch_stride; r.C_gs = p.C_group_stride; r.C_ns = p.C_dstate_stride;
    grid = di.sms * (per_sm > 4 ? 4 : per_sm);
    if (grid > r.n_units) grid = r.n_units;
    return true;
}

// Row-serial backward (scan_bwd_rows.cuh): eligibility + argument block.  Returns false when the warp-scan kernels must run.
bool plan_rows_bwd(const mia_ss_params &p, const DeviceInfo &di, mia::RowsBwdArgs &r, int &grid) {
    const int es = esize(p.itype), eo = esize(p.otype), L = p.seqlen;
    const int rpg = p.dim / p.n_groups;
    if (p.dstate != 1 || p.z || p.delta_dim != p.dim || (rpg % 32) || (L % 4)) return false;
    if (dbg_knob("MIA_NO_ROWS_BWD")) return false;
    const int CH = mia::kRowsChunk;
    const int nch = (L + CH - 1) / CH;
    // long rows: x must hold a checkpoint every 256 tokens, and the per-row pieces of a tile must be 16-byte aligned
    if (nch > 1 && (mia_ss_chunk_len(L) != CH || !p.x || ((L * es) % 16) || ((L * eo) % 16))) return false;
    auto dense = [&](long long bs, long long ds) { return ds == L && bs == (long long)p.dim * L; };
    if (!dense(p.u_batch_stride, p.u_d_stride) || !dense(p.delta_batch_stride, p.delta_d_stride) ||
        !dense(p.dout_batch_stride, p.dout_d_stride) || !dense(p.du_batch_stride, p.du_d_stride) ||
        !dense(p.ddelta_batch_stride, p.ddelta_d_stride)) return false;
    if (p.A_d_stride != 1 && p.dim > 1) return false;
    if (((uintptr_t)p.u | (uintptr_t)p.delta | (uintptr_t)p.dout | (uintptr_t)p.du | (uintptr_t)p.ddelta) & 15) return false;
    memset(&r, 0, sizeof(r));
    r.batch = p.batch; r.dim = p.dim; r.L = L; r.G = p.n_groups; r.rows_per_group = rpg; r.softplus = p.delta_softplus;
    r.n_chunks = nch;
    const int span = nch == 1 ? L : CH;                         // tokens of a row resident at a time
    r.t0_pct = dbg_int("MIA_T0_PCT", 52);   // share of a chunk for the half that has no Gs to accumulate
    const int T0 = (span * r.t0_pct / 100) / 4 * 4;
    const int longer = T0 > span - T0 ? T0 : span - T0;
    r.nblk = (longer + mia::kBlk - 1) / mia::kBlk;
    r.Lp = (span + mia::kBlk - 1) / mia::kBlk * mia::kBlk + mia::kBlk;
    r.n_items = p.batch * p.n_groups * (rpg / 32);
    r.tile_bytes = round_up(32 * (span * es + (nch > 1 ? 16 : 0)), 128);
    r.tileo_bytes = round_up(32 * (span * eo + (nch > 1 ? 16 : 0)), 128);
    r.off_delta = r.tile_bytes;
    r.off_dout = 2 * r.tile_bytes;
    r.off_bc32 = r.off_dout + r.tileo_bytes;
    r.off_ck = r.off_bc32 + round_up(2 * r.Lp * 4, 128);
    r.off_xch = r.off_ck + 3 * r.nblk * 128;                    // ck of both halves + ckm
    r.off_bar = r.off_xch + 7 * 128;
    r.smem_bytes = r.off_bar + 128;
    int per_sm = (227 * 1024) / (r.smem_bytes + 1024);
    if (per_sm < 3) return false;                               // too few resident warps to hide the dependent chains
    if (per_sm > 5) per_sm = 5;                                 // 64 threads x 168 registers: 3 warps per scheduler
    r.u = p.u; r.delta = p.delta; r.A = p.A; r.B = p.B; r.C = p.C; r.D = p.D; r.delta_bias = p.delta_bias; r.dout = p.dout;
    r.du = p.du; r.ddelta = p.ddelta; r.x = p.x;
    r.B_bs = p.B_batch_stride; r.B_gs = p.B_group_stride; r.C_bs = p.C_batch_stride; r.C_gs = p.C_group_stride;
    grid = di.sms * per_sm;
    // Long rows: a CTA walks its 32 rows chunk by chunk and each chunk is latency-bound (load -> phase 1 -> exchange ->
    // phase 2 -> store), so this path only wins while every item gets its own resident CTA (measured on B200, L = 6400:
    // 546 vs 734 us at 0.65 items per slot, 1076 vs 1123 us at 1.3, 2017 vs 1769 us at 2.6).
    if (nch > 1 && r.n_items > grid) return false;
    if (grid > r.n_items) grid = r.n_items;
    return true;
}

// Column-walk backward (scan_bwd_cw.cuh): eligibility, argument block, tensor maps (u, delta, dout, du, ddelta).
bool plan_cw_bwd(const mia_ss_params &p, const DeviceInfo &di, mia::CwBwdArgs &r, CUtensorMap *tm, int &grid) {
    const int es = esize(p.itype), eo = esize(p.otype), L = p.seqlen;
    int g = 0;
    if (!p.hblk || !cw_geometry(p, di.sms, g)) return false;
    if (dbg_knob("MIA_NO_CW_BWD")) return false;
    const int rpg = p.dim / p.n_groups;
    auto dense = [&](long long bs, long long ds) { return ds == L && bs == (long long)p.dim * L; };
    if (!dense(p.u_batch_stride, p.u_d_stride) || !dense(p.delta_batch_stride, p.delta_d_stride) ||
        !dense(p.dout_batch_stride, p.dout_d_stride) || !dense(p.du_batch_stride, p.du_d_stride) ||
        !dense(p.ddelta_batch_stride, p.ddelta_d_stride)) return false;
    if (p.A_d_stride != 1 && p.dim > 1) return false;
    if (((uintptr_t)p.u | (uintptr_t)p.delta | (uintptr_t)p.dout | (uintptr_t)p.du | (uintptr_t)p.ddelta) & 15) return false;
    const unsigned long long trows = (unsigned long long)p.batch * p.dim / g, tcols = (unsigned long long)g * L;
    if (tcols >= (1ull << 31) || trows >= (1ull << 31)) return false;
    memset(&r, 0, sizeof(r));
    r.batch = p.batch; r.dim = p.dim; r.L = L; r.G = p.n_groups; r.rows_per_group = rpg; r.softplus = p.delta_softplus;
    r.g = g;
    r.n_items = p.batch * p.n_groups * (rpg / (32 * g));
    r.ngrp = (g * L + mia::kCwGrp - 1) / mia::kCwGrp;
    r.nwin = (g * L + mia::kCwWin - 1) / mia::kCwWin;
    const int tile_i = 32 * mia::kCwWin * es, tile_o = 32 * mia::kCwWin * eo;
    r.stage_bytes = 2 * tile_i + tile_o;
    // Many short rows (at least two rounds of items at 8 per SM) want 8 resident warps per SM (see below) and can afford 3
    // stages; otherwise 2 stages with the half-step-early refill, so that 12 warps per SM fit
    const bool many_short = L <= 256 && r.n_items >= 2LL * 8 * di.sms;
    r.ns = dbg_int("MIA_CW_STAGES", many_short ? 3 : 2);
    r.off_bc32 = r.ns * r.stage_bytes;
    r.off_pf = r.off_bc32 + 2 * mia::kCwWin * 4;                // two 512-byte prefetch slots (raw B, raw C, block states of 2 groups)
    r.off_red = r.off_pf + 1024;                                // [32][20] fp32 scratch of the row reductions
    r.off_bar = r.off_red + 32 * 20 * 4;
    r.smem_bytes = r.off_bar + 8 * r.ns + 16;
    if ((((uintptr_t)p.B | (uintptr_t)p.C) & 3) || ((p.B_batch_stride | p.B_group_stride | p.C_batch_stride | p.C_group_stride) * es) % 4) return false;
    r.A = p.A; r.B = p.B; r.C = p.C; r.D = p.D; r.delta_bias = p.delta_bias; r.hblk = p.hblk;
    r.B_bs = p.B_batch_stride; r.B_gs = p.B_group_stride; r.C_bs = p.C_batch_stride; r.C_gs = p.C_group_stride;
    const void *ptrs[5] = {p.u, p.delta, p.dout, p.du, p.ddelta};
    for (int i = 0; i < 5; ++i) {
        const int e = i == 2 ? eo : es;
        const int trc = mia::tma_make_2d(&tm[i], ptrs[i], trows, tcols, tcols * e, 32, mia::kCwWin, e, mia::kCwWin * e);
        if (trc != 0) { MIA_TRACE("cw bwd: tensor map %d rejected (CUresult %d)", i, trc); return false; }
    }
    int per_sm = (227 * 1024) / (r.smem_bytes + 1024);
    if (per_sm > 12) per_sm = 12;                               // 32 threads x 168 registers: 3 warps per scheduler
    // Many short items: 8 resident warps per SM finish an item 1.8x faster than 12 do (measured, gpurun r2s / r2t, bf16,
    // B = 148: L = 196 0.313 against 0.354 ms, L = 200 0.302 / 0.362, L = 100 0.173 / 0.193, L = 104 0.173 / 0.184), so with
    // at least two rounds of short rows the extra rounds cost less than the contention.  Long rows (L = 1024, B = 32: 0.341
    // against 0.291 ms; L = 6400) want every slot.
    if (many_short && per_sm > 8) per_sm = 8;
    per_sm = std::min(per_sm, dbg_int("MIA_CW_MAXPERSM", per_sm));
    if (per_sm < 1) return false;
    // at most 8 resident warps per SM: the build that may use 255 registers (scan_bwd_cw.cuh, kMinBlk); same results
    r.wide = dbg_int("MIA_CW_WIDE", per_sm <= 8 ? 1 : 0);
    if (r.wide && per_sm > 8) per_sm = 8;
    const long long slots = (long long)di.sms * per_sm;
    const long long rounds = (r.n_items + slots - 1) / slots;
    grid = (int)((r.n_items + rounds - 1) / rounds);             // equal rounds per CTA
    return true;
}

// Deterministic d_state 16 backward (scan_bwd_rowsn.cuh): eligibility + argument block.
bool plan_rowsn_bwd(const mia_ss_params &p, const DeviceInfo &di, mia::RowsNBwdArgs &r, int &grid) {
    const int es = esize(p.itype), eo = esize(p.otype), L = p.seqlen;
    const int rpg = p.dim / p.n_groups;
    RowsNSplit sp;
    if (!rowsn_bwd_split(p, di.sms, kRowsNCtasPerSm, sp)) return false;
    if (dbg_knob("MIA_NO_ROWS_BWD")) return false;
    auto dense = [&](long long bs, long long ds) { return ds == L && bs == (long long)p.dim * L; };
    if (!dense(p.u_batch_stride, p.u_d_stride) || !dense(p.delta_batch_stride, p.delta_d_stride) ||
        !dense(p.dout_batch_stride, p.dout_d_stride) || !dense(p.du_batch_stride, p.du_d_stride) ||
        !dense(p.ddelta_batch_stride, p.ddelta_d_stride)) return false;
    if (((uintptr_t)p.u | (uintptr_t)p.delta | (uintptr_t)p.dout) & 15) return false;
    if (p.z) {
        if (!dense(p.z_batch_stride, p.z_d_stride) || !dense(p.out_saved_batch_stride, p.out_saved_d_stride) ||
            !dense(p.dz_batch_stride, p.dz_d_stride)) return false;
        if (((uintptr_t)p.z | (uintptr_t)p.out_saved) & 15) return false;
    }
    memset(&r, 0, sizeof(r));
    r.batch = p.batch; r.dim = p.dim; r.L = L; r.G = p.n_groups; r.rows_per_group = rpg;
    r.softplus = p.delta_softplus; r.has_z = p.z != nullptr;
    r.n_oct = sp.n_oct; r.oct_per_cta = sp.oct_per_cta; r.oct_per_group = sp.oct_per_group; r.max_parts = sp.max_parts;
    r.nblk = (L + mia::kRnBlk - 1) / mia::kRnBlk;
    r.Lp = r.nblk * mia::kRnBlk;
    r.pitch_s = r.Lp + 4;
    r.pitch_acc = r.Lp + 4;
    // B / C rows [16 states][pitch]: consecutive state rows 16 bytes apart modulo 128, so that the 8 state pairs of a
    // quarter warp read 8 different 16-byte bank windows
    r.pitch_bc = r.Lp;
    while ((r.pitch_bc * es) % 128 != 16) r.pitch_bc += 16 / es;
    int off = 0;
    auto take = [&](int bytes) { int o = off; off += round_up(bytes, 128); return o; };
    r.off_raw_u = take(mia::kRnOct * L * es);
    r.off_raw_d = take(mia::kRnOct * L * es);
    r.off_raw_o = take(mia::kRnOct * L * eo);
    if (p.z) { r.off_raw_z = take(mia::kRnOct * L * es); r.off_raw_os = take(mia::kRnOct * L * eo); }
    r.off_sm = take(4 * r.pitch_s * 4);
    r.off_su = take(4 * r.pitch_s * 4);
    r.off_sy = take(4 * r.pitch_s * 4);
    r.off_b = take(16 * r.pitch_bc * es);
    r.off_c = take(16 * r.pitch_bc * es);
    r.off_acc = take(2 * 16 * r.pitch_acc * 4);
    r.off_ck = take(r.nblk * 32 * 8);
    r.off_ckm = take(r.nblk * 4 * 4);
    r.off_x = take((4 * mia::kRnWarps * 32 + mia::kRnWarps * 4) * 8);
    r.off_bar = take(8);
    r.smem_bytes = off;
    if (r.smem_bytes > di.smem_optin) return false;
    r.u = p.u; r.delta = p.delta; r.A = p.A; r.B = p.B; r.C = p.C; r.D = p.D; r.delta_bias = p.delta_bias; r.dout = p.dout;
    r.z = p.z; r.out_saved = p.out_saved; r.du = p.du; r.ddelta = p.ddelta; r.dz = p.dz;
    r.A_ds = p.A_d_stride; r.A_ns = p.A_dstate_stride;
    r.B_bs = p.B_batch_stride; r.B_gs = p.B_group_stride; r.B_ns = p.B_dstate_stride;
    r.C_bs = p.C_batch_stride; r.C_gs = p.C_group_stride; r.C_ns = p.C_dstate_stride;
    grid = sp.grid;
    return true;
}

template <typename F>
int dispatch(int itype, F &&f) {
    switch (itype) {
        case MIA_F32: return f((float *)nullptr);
        case MIA_F16: return f((__half *)nullptr);
        default: return f((__nv_bfloat16 *)nullptr);
    }
}

// ---------------------------------------------------------------------------------------------
// Finalize: fold the backward's partials in a fixed order (deterministic) and cast to the output dtypes.
struct FinArgs {
    int batch, dim, L, N, G, delta_dim, ratio, tiles, bc_atomic, Lp, has_D, has_bias;
    int oct_per_cta, oct_per_group;          // > 0: partials of scan_bwd_rowsn.cuh, group bg has (c_hi - c_lo + 1) <= tiles of them
    const float *part_dA, *part_dD, *part_dbias, *acc_dB, *acc_dC, *ddelta_full;
    float *dA, *dD, *dbias;
    void *dB, *dC, *ddelta;
    long long dA_ds, dA_ns, dB_bs, dB_gs, dB_ns, dC_bs, dC_gs, dC_ns, dd_bs, dd_ds;
    long long n_bc, n_dA, n_dD, n_db, n_dd;
};

// thread items: dB, dC (sum of the per-segment partials / cast of the atomic accumulator) and the ddelta group fold;
// warp items: dA, dD, dbias (sum over the batch, lanes stride over b, shuffle-reduced in a fixed order).
template <typename T>
__global__ void ss_finalize_kernel(const __grid_constant__ FinArgs f) {
    using raw = typename mia::Cvt<T>::raw;
    const long long n_thread_items = 2 * f.n_bc + f.n_dd;
    const long long t1 = (n_thread_items + 31) / 32 * 32;
    const long long n_warp_items = f.n_dA + f.n_dD + f.n_db;
    const long long total = t1 + 32 * n_warp_items;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        if (idx < t1) {
            long long i = idx;
            if (i >= n_thread_items) continue;
            if (i < 2 * f.n_bc) {
                const bool isC = i >= f.n_bc;
                if (isC) i -= f.n_bc;
                const int l = (int)(i % f.L);
                long long t = i / f.L;
                const int n = (int)(t % f.N); t /= f.N;
                const int g = (int)(t % f.G);
                const int b = (int)(t / f.G);
                const float *acc = isC ? f.acc_dC : f.acc_dB;
                float sum = 0.f;
                if (f.bc_atomic) {
                    sum = acc[((size_t)(b * f.G + g) * f.N + n) * f.Lp + l];
                } else {
                    const float *p0 = acc + (((size_t)(b * f.G + g) * f.tiles) * f.N + n) * f.L + l;
                    const size_t step = (size_t)f.N * f.L;
                    int nt = f.tiles;
                    if (f.oct_per_cta > 0) {             // partial slots actually written for this group (CTAs it straddles)
                        const long long o0 = (long long)(b * f.G + g) * f.oct_per_group;
                        nt = (int)((o0 + f.oct_per_group - 1) / f.oct_per_cta - o0 / f.oct_per_cta) + 1;
                    }
#pragma unroll 4
                    for (int tl = 0; tl < nt; ++tl) sum += p0[tl * step];
                }
                raw *dst = reinterpret_cast<raw *>(isC ? f.dC : f.dB);
                const long long o = isC ? (b * f.dC_bs + g * f.dC_gs + n * f.dC_ns + l) : (b * f.dB_bs + g * f.dB_gs + n * f.dB_ns + l);
                dst[o] = mia::Cvt<T>::from_f(sum);
            } else {   // ddelta fold over the delta group (selective_scan_oflex.cpp:348-350)
                i -= 2 * f.n_bc;
                const int l = (int)(i % f.L);
                long long t = i / f.L;
                const int dg = (int)(t % f.delta_dim);
                const int b = (int)(t / f.delta_dim);
                float sum = 0.f;
                for (int r = 0; r < f.ratio; ++r) sum += f.ddelta_full[((size_t)b * f.dim + dg * f.ratio + r) * f.L + l];
                reinterpret_cast<raw *>(f.ddelta)[b * f.dd_bs + dg * f.dd_ds + l] = mia::Cvt<T>::from_f(sum);
            }
            continue;
        }
        // ---- warp items (idx - t1 is warp aligned, so the whole warp takes this path together)
        long long w = (idx - t1) >> 5;
        const int lane = (int)(idx & 31);
        float sum = 0.f;
        float *dst;
        if (w < f.n_dA) {
            const int n = (int)(w % f.N), d = (int)(w / f.N);
            for (int b = lane; b < f.batch; b += 32) sum += f.part_dA[((size_t)b * f.dim + d) * f.N + n];
            dst = f.dA + d * f.dA_ds + n * f.dA_ns;
        } else if (w < f.n_dA + f.n_dD) {
            w -= f.n_dA;
            for (int b = lane; b < f.batch; b += 32) sum += f.part_dD[(size_t)b * f.dim + w];
            dst = f.dD + w;
        } else {
            w -= f.n_dA + f.n_dD;
            for (int k = lane; k < f.batch * f.ratio; k += 32) {
                const int b = k / f.ratio, r = k - b * f.ratio;
                sum += f.part_dbias[(size_t)b * f.dim + w * f.ratio + r];
            }
            dst = f.dbias + w;
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
        if (lane == 0) *dst = sum;
    }
}

}  // namespace

// =================================================================================================
extern "C" {

int mia_abi_version(void) { return MIA_ABI_VERSION; }
const char *mia_last_error(void) { return g_err; }
uint64_t mia_launch_count(void) { return g_launches.load(); }

int mia_ss_chunk_len(int seqlen) { return lanes_per_row(seqlen) * mia::kTok; }
int mia_ss_num_chunks(int seqlen) {
    const int ch = mia_ss_chunk_len(seqlen);
    return (seqlen + ch - 1) / ch;
}

size_t mia_ss_block_state_floats(const mia_ss_params *pp) {
    if (!pp || validate_sizes(*pp) != MIA_OK) return 0;
    const mia_ss_params &p = *pp;
    if (p.dstate != 1 || ((p.dim / p.n_groups) % 32)) return 0;
    return (size_t)p.batch * p.dim * ((p.seqlen + 15) / 16);
}

int mia_ss_fwd_writes_block_states(const mia_ss_params *pp) {
    if (!pp || !pp->hblk || validate_sizes(*pp) != MIA_OK || !pp->u || !pp->delta || !pp->out) return 0;
    DeviceInfo di;
    if (device_info(di) != MIA_OK) return 0;
    mia::CwFwdArgs rw;
    CUtensorMap tm[3];
    int grid = 0;
    return plan_cw_fwd(*pp, di, rw, tm, grid) ? 1 : 0;          // the column-walk forward is the kernel that fills hblk
}

int mia_selective_scan_fwd(const mia_ss_params *pp, void *cuda_stream) {
    if (!pp) return fail(MIA_EINVAL, "null params");
    const mia_ss_params &p = *pp;
    if (int rc = validate_common(p)) return rc;
    MIA_CHECK(p.out && p.x, "out and x must not be null");
    MIA_CHECK(!p.z || p.out_z, "out_z is required when z is given");
    DeviceInfo di;
    if (int rc = device_info(di)) return rc;
    cudaStream_t stream = (cudaStream_t)cuda_stream;
    {
        mia::CwFwdArgs r;
        CUtensorMap tm[3];
        int rgrid = 0;
        if (plan_cw_fwd(p, di, r, tm, rgrid)) {
            MIA_TRACE("fwd: cw g=%d ns=%d nwin=%d grid=%d smem=%d", r.g, r.ns, r.nwin, rgrid, r.smem_bytes);
            const bool of32 = p.otype == MIA_F32 && p.itype != MIA_F32;
            const int rc = dispatch(p.itype, [&](auto *tag) {
                using T = typename std::remove_pointer<decltype(tag)>::type;
                return (int)mia::launch_fwd_cw<T>(tm, r, rgrid, of32, stream);
            });
            if (rc != 0) return fail(MIA_ECUDA, "selective_scan_fwd (column-walk) launch: %s", cudaGetErrorString((cudaError_t)rc));
            g_launches.fetch_add(1);
            return MIA_OK;
        }
    }
    {
        mia::ChunkArgs r;
        int rgrid = 0;
        if (plan_chunks_fwd(p, di, r, rgrid)) {
            const bool of32 = p.otype == MIA_F32 && p.itype != MIA_F32;
            const int rc = dispatch(p.itype, [&](auto *tag) {
                using T = typename std::remove_pointer<decltype(tag)>::type;
                return (int)mia::launch_fwd_chunks<T>(r, rgrid, of32, stream);
            });
            if (rc != 0) return fail(MIA_ECUDA, "selective_scan_fwd (chunk-parallel) launch: %s", cudaGetErrorString((cudaError_t)rc));
            g_launches.fetch_add(3);
            return MIA_OK;
        }
    }
    {
        mia::RowsArgs r;
        int rgrid = 0;
        if (plan_rows_fwd(p, di, r, rgrid)) {
            const bool of32 = p.otype == MIA_F32 && p.itype != MIA_F32;
            const int rc = dispatch(p.itype, [&](auto *tag) {
                using T = typename std::remove_pointer<decltype(tag)>::type;
                return (int)mia::launch_fwd_rows<T>(r, rgrid, of32, stream);
            });
            if (rc != 0) return fail(MIA_ECUDA, "selective_scan_fwd (row-serial) launch: %s", cudaGetErrorString((cudaError_t)rc));
            g_launches.fetch_add(1);
            return MIA_OK;
        }
    }
    {
        mia::StreamArgs r;
        int rgrid = 0;
        if (plan_stream_fwd(p, di, r, rgrid)) {
            const bool of32 = p.otype == MIA_F32 && p.itype != MIA_F32;
            const int rc = dispatch(p.itype, [&](auto *tag) {
                using T = typename std::remove_pointer<decltype(tag)>::type;
                return (int)mia::launch_fwd_stream<T>(r, rgrid, of32, stream);
            });
            if (rc != 0) return fail(MIA_ECUDA, "selective_scan_fwd (streaming) launch: %s", cudaGetErrorString((cudaError_t)rc));
            g_launches.fetch_add(1);
            return MIA_OK;
        }
    }
    {
        mia::RowsNArgs r;
        int rgrid = 0;
        if (plan_rowsn_fwd(p, di, r, rgrid)) {
            const bool of32 = p.otype == MIA_F32 && p.itype != MIA_F32;
            const int rc = dispatch(p.itype, [&](auto *tag) {
                using T = typename std::remove_pointer<decltype(tag)>::type;
                return (int)mia::launch_fwd_rowsn<T>(r, rgrid, of32, stream);
            });
            if (rc != 0) return fail(MIA_ECUDA, "selective_scan_fwd (row-serial, d_state %d) launch: %s", p.dstate, cudaGetErrorString((cudaError_t)rc));
            g_launches.fetch_add(1);
            return MIA_OK;
        }
    }
    Plan pl;
    mia::ScanArgs a;
    if (int rc = plan_and_layout(p, di, false, pl, a)) return rc;
    a.out = p.out; a.out_z = p.out_z;
    a.out_bs = p.out_batch_stride; a.out_ds = p.out_d_stride; a.outz_bs = p.out_z_batch_stride; a.outz_ds = p.out_z_d_stride;
    const int grid = a.n_seg < di.sms ? a.n_seg : di.sms;
    const int rc = dispatch(p.itype, [&](auto *tag) {
        using T = typename std::remove_pointer<decltype(tag)>::type;
        return (int)mia::launch_fwd_any<T>(a, grid, stream);
    });
    if (rc != 0) return fail(MIA_ECUDA, "selective_scan_fwd launch: %s", cudaGetErrorString((cudaError_t)rc));
    g_launches.fetch_add(1);
    return MIA_OK;
}

size_t mia_selective_scan_bwd_workspace(const mia_ss_params *pp) {
    if (!pp) return 0;
    if (validate_sizes(*pp) != MIA_OK) return 0;     // sizes that the backward itself would reject
    DeviceInfo di;
    if (device_info(di) != MIA_OK) { di.sms = 148; di.smem_optin = 232448; }
    Plan pl;
    mia::ScanArgs a;
    if (plan_and_layout(*pp, di, true, pl, a) != MIA_OK) return 0;
    return workspace_layout(*pp, pl, di.sms).total;
}

int mia_selective_scan_bwd(const mia_ss_params *pp, void *cuda_stream) {
    if (!pp) return fail(MIA_EINVAL, "null params");
    const mia_ss_params &p = *pp;
    if (int rc = validate_common(p)) return rc;
    MIA_CHECK(p.dout && p.du && p.ddelta && p.dA && p.dB && p.dC, "dout, du, ddelta, dA, dB, dC must not be null");
    MIA_CHECK(!p.D || p.dD, "dD is required when D is given");
    MIA_CHECK(!p.delta_bias || p.ddelta_bias, "ddelta_bias is required when delta_bias is given");
    MIA_CHECK(!p.z || (p.dz && p.out_saved), "dz and out_saved are required when z is given");
    MIA_CHECK(p.n_chunks == 1 || p.x, "x is required when the sequence spans more than one chunk");
    DeviceInfo di;
    if (int rc = device_info(di)) return rc;
    Plan pl;
    mia::ScanArgs a;
    if (int rc = plan_and_layout(p, di, true, pl, a)) return rc;
    a.dout = p.dout; a.out_saved = p.out_saved; a.du = p.du; a.ddelta = p.ddelta; a.dz = p.dz;
    a.dout_bs = p.dout_batch_stride; a.dout_ds = p.dout_d_stride;
    a.osaved_bs = p.out_saved_batch_stride; a.osaved_ds = p.out_saved_d_stride;
    a.du_bs = p.du_batch_stride; a.du_ds = p.du_d_stride; a.dd_bs = p.ddelta_batch_stride; a.dd_ds = p.ddelta_d_stride;
    a.dz_bs = p.dz_batch_stride; a.dz_ds = p.dz_d_stride;
    const WorkspaceLayout w = workspace_layout(p, pl, di.sms);
    if (!p.workspace || p.workspace_bytes < w.total)
        return fail(MIA_EWORKSPACE, "workspace too small: need %zu bytes, got %zu", w.total, p.workspace ? p.workspace_bytes : (size_t)0);
    MIA_CHECK(((uintptr_t)p.workspace & 255) == 0, "workspace must be 256-byte aligned");
    char *ws = (char *)p.workspace;
    a.part_dA = (float *)(ws + w.part_dA); a.part_dD = (float *)(ws + w.part_dD); a.part_dbias = (float *)(ws + w.part_dbias);
    a.acc_dB = (float *)(ws + w.acc_dB); a.acc_dC = (float *)(ws + w.acc_dC);
    a.ddelta_full = (float *)(ws + w.ddelta_full);
    a.bc_atomic = w.bc_atomic;
    cudaStream_t stream = (cudaStream_t)cuda_stream;
    mia::RowsNBwdArgs rn;
    int rn_grid = 0;
    const bool use_rowsn = plan_rowsn_bwd(p, di, rn, rn_grid);
    if (use_rowsn) a.bc_atomic = 0;
    if (a.bc_atomic) {
        MIA_CUDA(cudaMemsetAsync(a.acc_dB, 0, w.acc_bytes, stream));
        MIA_CUDA(cudaMemsetAsync(a.acc_dC, 0, w.acc_bytes, stream));
        g_launches.fetch_add(2);
    }
    int rc = 0, bc_parts = pl.split;
    mia::RowsBwdArgs rb;
    int rgrid = 0;
    CUtensorMap tmaps[5];
    // Column-walk kernel (one warp per 32 g rows, phase 2 only, on the forward's block states): short rows always; rows of more
    // than one 256-token chunk only with enough 32-row items to fill the SMs.  Measured (gpurun r2i / r2n, L = 6400, bf16):
    // B = 16 (1536 items) 0.86 ms against 1.77 ms for the warp-scan kernel; B = 4 (384 items = 2.6 warps per SM): a one-warp-
    // per-item walk took 0.75 ms against 0.55 ms.
    const long long items32 = (long long)p.batch * (p.dim / 32);
    mia::CwBwdArgs cwb;
    const bool want_cw = !use_rowsn && p.hblk && (p.seqlen <= mia::kRowsChunk || items32 >= 4LL * di.sms || dbg_knob("MIA_FORCE_CW_BWD"));
    if (want_cw && plan_cw_bwd(p, di, cwb, tmaps, rgrid)) {
        MIA_TRACE("bwd: cw g=%d ns=%d ngrp=%d grid=%d smem=%d", cwb.g, cwb.ns, cwb.ngrp, rgrid, cwb.smem_bytes);
        cwb.part_dA = a.part_dA; cwb.part_dD = a.part_dD; cwb.part_dbias = a.part_dbias; cwb.acc_dB = a.acc_dB; cwb.acc_dC = a.acc_dC;
        bc_parts = cwb.rows_per_group / 32;
        const bool of32 = p.otype == MIA_F32 && p.itype != MIA_F32;
        rc = dispatch(p.itype, [&](auto *tag) {
            using T = typename std::remove_pointer<decltype(tag)>::type;
            return (int)mia::launch_bwd_cw<T>(tmaps, cwb, rgrid, of32, stream);
        });
    } else if (use_rowsn) {
        rn.part_dA = a.part_dA; rn.part_dD = a.part_dD; rn.part_dbias = a.part_dbias; rn.acc_dB = a.acc_dB; rn.acc_dC = a.acc_dC;
        bc_parts = rn.max_parts;
        const bool of32 = p.otype == MIA_F32 && p.itype != MIA_F32;
        rc = dispatch(p.itype, [&](auto *tag) {
            using T = typename std::remove_pointer<decltype(tag)>::type;
            return (int)mia::launch_bwd_rowsn<T>(rn, rn_grid, of32, stream);
        });
    } else if (plan_rows_bwd(p, di, rb, rgrid)) {
        rb.part_dA = a.part_dA; rb.part_dD = a.part_dD; rb.part_dbias = a.part_dbias; rb.acc_dB = a.acc_dB; rb.acc_dC = a.acc_dC;
        bc_parts = rb.rows_per_group / 32;
        const bool of32 = p.otype == MIA_F32 && p.itype != MIA_F32;
        rc = dispatch(p.itype, [&](auto *tag) {
            using T = typename std::remove_pointer<decltype(tag)>::type;
            return (int)mia::launch_bwd_rows<T>(rb, rgrid, of32, stream);
        });
    } else {
        const int grid = a.n_seg < di.sms ? a.n_seg : di.sms;
        rc = dispatch(p.itype, [&](auto *tag) {
            using T = typename std::remove_pointer<decltype(tag)>::type;
            return (int)mia::launch_bwd_any<T>(a, grid, stream);
        });
    }
    if (rc != 0) return fail(MIA_ECUDA, "selective_scan_bwd launch: %s", cudaGetErrorString((cudaError_t)rc));
    g_launches.fetch_add(1);

    FinArgs f;
    memset(&f, 0, sizeof(f));
    f.batch = p.batch; f.dim = p.dim; f.L = p.seqlen; f.N = p.dstate; f.G = p.n_groups; f.delta_dim = p.delta_dim;
    f.ratio = p.dim / p.delta_dim; f.tiles = bc_parts; f.bc_atomic = a.bc_atomic; f.Lp = (p.seqlen + 3) & ~3;
    if (use_rowsn) { f.oct_per_cta = rn.oct_per_cta; f.oct_per_group = rn.oct_per_group; }
    f.part_dA = a.part_dA; f.part_dD = a.part_dD; f.part_dbias = a.part_dbias; f.acc_dB = a.acc_dB; f.acc_dC = a.acc_dC;
    f.ddelta_full = a.ddelta_full;
    f.dA = p.dA; f.dD = p.dD; f.dbias = p.ddelta_bias; f.dB = p.dB; f.dC = p.dC; f.ddelta = p.ddelta;
    f.dA_ds = p.dA_d_stride; f.dA_ns = p.dA_dstate_stride;
    f.dB_bs = p.dB_batch_stride; f.dB_gs = p.dB_group_stride; f.dB_ns = p.dB_dstate_stride;
    f.dC_bs = p.dC_batch_stride; f.dC_gs = p.dC_group_stride; f.dC_ns = p.dC_dstate_stride;
    f.dd_bs = p.ddelta_batch_stride; f.dd_ds = p.ddelta_d_stride;
    f.n_bc = (long long)p.batch * p.n_groups * p.dstate * p.seqlen;
    f.n_dA = (long long)p.dim * p.dstate;
    f.n_dD = p.D ? p.dim : 0;
    f.n_db = p.delta_bias ? p.delta_dim : 0;
    f.n_dd = f.ratio > 1 ? (long long)p.batch * p.delta_dim * p.seqlen : 0;
    const long long total = (2 * f.n_bc + f.n_dd + 31) / 32 * 32 + 32 * (f.n_dA + f.n_dD + f.n_db);
    long long blocks = (total + 255) / 256;
    if (blocks > di.sms * 8) blocks = di.sms * 8;
    rc = dispatch(p.itype, [&](auto *tag) {
        using T = typename std::remove_pointer<decltype(tag)>::type;
        ss_finalize_kernel<T><<<(int)blocks, 256, 0, stream>>>(f);
        return (int)cudaGetLastError();
    });
    if (rc != 0) return fail(MIA_ECUDA, "selective_scan_bwd finalize launch: %s", cudaGetErrorString((cudaError_t)rc));
    g_launches.fetch_add(1);
    return MIA_OK;
}

}  // extern "C"
