// engine.cpp — adjustment context and the C-ABI (include/gadj.h).
//
// Host orchestration of one Gauss–Newton iteration on the device:
//   assemble (N, w)  ->  equilibrate + scatter into front panels  ->  supernodal
//   Cholesky  ->  forward/backward solve  ->  x += delta  ->  [selected inverse]
// mirroring dna_adjust::PrepareAdjustment / AdjustSimultaneous / Solve /
// GenerateStatistics (ADJ:258, ADJ:2413-2511, ADJ:6586-6667, ADJ:6802-6841).
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/gadj.h"
#include "dev.h"
#include "geodesy.h"
#include "kernels.h"
#include "par.h"
#include "plan.h"
#include "rows.h"
#include "symbolic.h"

using namespace gadj;

namespace {

std::string g_create_error;

// acklam + one Halley step; stands in for boost::math::quantile(normal) (ADJ:203-206)
double norm_quantile(double p)
{
    static const double a[] = {-3.969683028665376e+01, 2.209460984245205e+02, -2.759285104469687e+02,
                               1.383577518672690e+02,  -3.066479806614716e+01, 2.506628277459239e+00};
    static const double b[] = {-5.447609879822406e+01, 1.615858368580409e+02, -1.556989798598866e+02,
                               6.680131188771972e+01,  -1.328068155288572e+01};
    static const double c[] = {-7.784894002430293e-03, -3.223964580411365e-01, -2.400758277161838e+00,
                               -2.549732539343734e+00, 4.374664141464968e+00,  2.938163982698783e+00};
    static const double d[] = {7.784695709041462e-03, 3.224671290700398e-01, 2.445134137142996e+00,
                               3.754408661907416e+00};
    double q, r, x;
    if (p < 0.02425) {
        q = std::sqrt(-2 * std::log(p));
        x = (((((c[0] * q + c[1]) * q + c[2]) * q + c[3]) * q + c[4]) * q + c[5]) /
            ((((d[0] * q + d[1]) * q + d[2]) * q + d[3]) * q + 1);
    } else if (p <= 1 - 0.02425) {
        q = p - 0.5;
        r = q * q;
        x = (((((a[0] * r + a[1]) * r + a[2]) * r + a[3]) * r + a[4]) * r + a[5]) * q /
            (((((b[0] * r + b[1]) * r + b[2]) * r + b[3]) * r + b[4]) * r + 1);
    } else {
        q = std::sqrt(-2 * std::log(1 - p));
        x = -(((((c[0] * q + c[1]) * q + c[2]) * q + c[3]) * q + c[4]) * q + c[5]) /
            ((((d[0] * q + d[1]) * q + d[2]) * q + d[3]) * q + 1);
    }
    double e = 0.5 * std::erfc(-x / std::sqrt(2.0)) - p;
    double u = e * std::sqrt(2 * kPi) * std::exp(x * x / 2);
    return x - u / (1 + x * u / 2);
}

void mat3_mul(const double* A, bool tA, const double* B, bool tB, double* C)  // row-major
{
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double s = 0;
            for (int k = 0; k < 3; ++k)
                s += (tA ? A[k * 3 + i] : A[i * 3 + k]) * (tB ? B[j * 3 + k] : B[k * 3 + j]);
            C[i * 3 + j] = s;
        }
}

bool mat3_inverse(const double* A, double* B)
{
    double det = A[0] * (A[4] * A[8] - A[5] * A[7]) - A[1] * (A[3] * A[8] - A[5] * A[6]) + A[2] * (A[3] * A[7] - A[4] * A[6]);
    if (det == 0.0 || std::isnan(det))
        return false;
    double id = 1.0 / det;
    B[0] = (A[4] * A[8] - A[5] * A[7]) * id;
    B[1] = (A[2] * A[7] - A[1] * A[8]) * id;
    B[2] = (A[1] * A[5] - A[2] * A[4]) * id;
    B[3] = (A[5] * A[6] - A[3] * A[8]) * id;
    B[4] = (A[0] * A[8] - A[2] * A[6]) * id;
    B[5] = (A[2] * A[3] - A[0] * A[5]) * id;
    B[6] = (A[3] * A[7] - A[4] * A[6]) * id;
    B[7] = (A[1] * A[6] - A[0] * A[7]) * id;
    B[8] = (A[0] * A[4] - A[1] * A[3]) * id;
    return true;
}

// geographic -> Cartesian Jacobian (dnatemplatematrixfuncs.hpp:204-232), row-major
void cart_geo_jacobian(const Ellipsoid& e, double lat, double lon, double h, double* R)
{
    double cl = std::cos(lat), sl = std::sin(lat), co = std::cos(lon), so = std::sin(lon);
    double t1a = e.a * e.e2;
    double ome = 1. - e.e2;
    double nu = prime_vertical(e, lat);
    double nuh = nu + h;
    double nu1h = nu * ome + h;
    double t1b = t1a * sl * cl;
    double t1c = std::pow((1. - e.e2 * sl * sl), 1.5);
    R[0] = (t1b * cl * co / t1c) - (nuh * sl * co);
    R[1] = -nuh * cl * so;
    R[2] = cl * co;
    R[3] = (t1b * cl * so / t1c) - (nuh * sl * so);
    R[4] = nuh * cl * co;
    R[5] = cl * so;
    R[6] = (t1b * ome * sl / t1c) + (nu1h * cl);
    R[7] = 0.;
    R[8] = sl;
}

template <class T>
struct DevArray {
    T* p = nullptr;
    size_t n = 0;
    bool shared = false;     // peer-visible allocation (multi-GPU exchange buffers)
    ~DevArray() { release(); }
    void release()
    {
        if (p) {
            if (shared)
                dev::free_shared(p);
            else
                dev::free_(p);
        }
        p = nullptr;
        n = 0;
    }
    bool resize(size_t count, bool peer_visible = false)
    {
        release();
        shared = peer_visible;
        if (count == 0)
            return true;
        p = (T*)(shared ? dev::alloc_shared(count * sizeof(T)) : dev::alloc(count * sizeof(T)));
        n = p ? count : 0;
        return p != nullptr;
    }
    bool upload(const std::vector<T>& v)
    {
        if (!resize(v.size()))
            return false;
        if (!v.empty())
            dev::h2d(p, v.data(), v.size() * sizeof(T));
        return true;
    }
    size_t bytes() const { return n * sizeof(T); }
};

}  // namespace

struct gadj_ctx {
    gadj_opts o{};
    std::string err;
    dev::Device* device = nullptr;   // this context's device state (ordinal, stream); made current by every entry point
    Ellipsoid ell{};
    // borrowed host data
    dna_stn_t* stn = nullptr;
    uint32_t nstn = 0;
    dna_msr_t* msr = nullptr;
    uint64_t nmsr = 0;
    int reduced = 0;
    std::vector<uint32_t> isl_off, isl;
    // measurement plan
    std::vector<uint32_t> first, edge_word;          // GNSS baselines
    std::vector<uint32_t> binc_ptr, binc;            // station -> incident GNSS baselines (bit 31: the station is station2)
    std::vector<uint32_t> edge_bsl;                  // per edge: its only baseline (block read from the baseline's slot) or ~0u
    // design rows of every other type (rows.h): terrestrial rows, derived angles of D sets, X / Y cluster rows
    std::vector<RowDesc> rows;
    std::vector<uint32_t> row_base;                  // D rows: record that supplies station2 / term3 / term4 of the angle
    std::vector<ClusterDesc> clusters;
    std::vector<uint32_t> cstn, inc_ptr, inc, pair_word;
    std::vector<double> cvmat;                       // cluster variance matrices (host copy, inverted on the device)
    uint64_t nrows = 0;
    bool non_gps = false;
    std::vector<uint32_t> edge_hi, edge_lo;
    bool contiguous = false;
    uint64_t nbsl = 0, nedge = 0;
    uint32_t constrained_components = 0, unused_stations = 0;
    // structure
    Symbolic sym;
    Plan plan;
    bool prepared = false, factor_valid = false, inverse_valid = false, normals_valid = false;
    bool stage_normals = false, vcv_extracted = false;
    int32_t mg_rank = 0, mg_world = 1;
    // multi-GPU: peer table (byte offsets to the other ranks' replicas, their barrier counters), filled by gadj_mg_connect
    bool connected = false;
    PeerTable peers{};
    DevArray<PeerTable> d_peers;
    DevArray<unsigned long long> d_counter;
    uint64_t barrier_round = 0;
    std::vector<std::pair<void*, int64_t>> peer_maps;   // mapped peer buffers (pointer, owning process) to unmap on destroy
    DevArray<ReduceOp> d_reduce, d_reduce_misc;
    DevArray<PushOp> d_push;
    DevArray<double*> d_bases;          // this rank's replicated buffers by McBuf index
    DevArray<double> d_apply;
    DevArray<uint8_t> d_pos_owned;
    uint32_t iteration = 0;
    double critical = 0;
    // device state
    DevArray<dna_msr_t> d_msr;
    DevArray<uint32_t> d_first, d_edge, d_edge_hi, d_edge_lo, d_pos, d_diag_ld, d_off_ld, d_binc_ptr, d_binc, d_edge_bsl;
    DevArray<double> d_bq;
    DevArray<RowDesc> d_rows;
    DevArray<ClusterDesc> d_clusters;
    DevArray<uint32_t> d_cstn, d_inc_ptr, d_inc, d_pair_word;
    DevArray<double> d_cvinv, d_row_l, d_row_a, d_row_t;
    DevArray<float> d_geoid;
    DevArray<uint64_t> d_diag_dest, d_off_dest;
    DevArray<double> d_est, d_est0, d_llh, d_llh0, d_cblock, d_ndiag, d_noff, d_w, d_dscale, d_panels, d_pool, d_wbuf, d_x, d_y, d_corr, d_vcvd,
        d_vcvo, d_sums;
    DevArray<int32_t> d_rowmap, d_rowidx, d_info, d_coltgt;
    DevArray<ScatterTarget> d_tgt;
    DevArray<GemmOp> d_gemm;
    DevArray<GemmTile> d_tiles;
    DevArray<DiagOp> d_diag;
    DevArray<TrimvOp> d_tri;
    DevArray<GemvOp> d_gemv;
    DevArray<TransposeOp> d_tr;
    DevArray<GatherOp> d_gather;
    DevArray<GatherTile> d_gather_tiles;
    const PeerTable* pt() const { return d_peers.p; }
    std::vector<double> h_corr, h_dscale;
    uint32_t h_dscale_iteration = ~0u;
    void* ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    uint64_t device_bytes = 0;
    // profiling
    bool profiling = false;
    struct ProfItem {
        int kind;
        double flops;
        int tiles;
        int tag, level;
    };
    std::vector<void*> prof_ev;          // two events per item
    std::vector<ProfItem> prof_items;
    gadj_profile prof{};
    uint64_t launch_count = 0;

    void prof_begin(int kind, double flops = 0, int tiles = 0, int tag = 0, int level = -1)
    {
        launch_count++;
        if (!profiling)
            return;
        size_t i = prof_items.size();
        while (prof_ev.size() < 2 * (i + 1))
            prof_ev.push_back(dev::event_create());
        prof_items.push_back({kind, flops, tiles, tag, level});
        dev::event_record(prof_ev[2 * i]);
    }
    void prof_end()
    {
        if (!profiling)
            return;
        dev::event_record(prof_ev[2 * (prof_items.size() - 1) + 1]);
    }

    int fail(const std::string& m)
    {
        err = m;
        return 1;
    }
};

namespace {

// all ranks meet on the device (no host synchronisation): see launch_barrier
void mg_barrier(gadj_ctx* c)
{
    if (c->mg_world <= 1)
        return;
    c->barrier_round++;
    launch_barrier(c->pt(), c->barrier_round * (uint64_t)c->mg_world, c->d_info.p, dev::stream());
}

void run_one(gadj_ctx* c, const Launch& L)
{
    void* st = dev::stream();
    {
        if (L.kind == L_BARRIER) {
            mg_barrier(c);
            return;
        }
        c->prof_begin(L.kind, L.flops, L.total_tiles, L.tag, L.level);
        if (L.kind == L_ZERO)
            c->launch_count--;  // a memset, not one of our kernels
        switch (L.kind) {
        case L_GEMM:
            launch_gemm(c->d_gemm.p + L.op_begin, L.op_count, c->d_tiles.p + L.tile_begin, L.total_tiles, L.shape, st);
            break;
        case L_DIAG:
            launch_diag(c->d_diag.p + L.op_begin, L.op_count, c->d_info.p, st);
            break;
        case L_TRI_FWD:
        case L_TRI_BWD:
            launch_trimv(c->d_tri.p + L.op_begin, L.op_count, st);
            break;
        case L_GEMV_FWD:
            launch_gemv(c->d_gemv.p + L.op_begin, L.op_count, c->d_x.p, c->d_x.p, 0, st);
            break;
        case L_GEMV_BWD:
            launch_gemv(c->d_gemv.p + L.op_begin, L.op_count, c->d_x.p, c->d_x.p, 1, st);
            break;
        case L_TRANSPOSE:
            launch_transpose(c->d_tr.p + L.op_begin, L.op_count, L.total_tiles, st);
            break;
        case L_GATHER:
            launch_gather(c->d_gather.p + L.op_begin, L.op_count, c->d_gather_tiles.p + L.tile_begin, L.total_tiles, st);
            break;
        case L_ALLREDUCE: {
            double* base = L.buf == MC_PANELS ? c->d_panels.p : L.buf == MC_X ? c->d_x.p : nullptr;
            if (base)
                launch_allreduce(c->d_reduce.p + L.op_begin, L.op_count, c->pt(), base, L.buf, st);
            break;
        }
        case L_PUSH:
            launch_push(c->d_push.p + L.op_begin, L.op_count, L.total_tiles, c->pt(), c->d_bases.p, st);
            break;
        case L_ZERO:
            dev::zero(L.zero_ptr, L.zero_bytes);
            break;
        }
        c->prof_end();
        // debugging aid (GADJ_SYNC_EVERY_LAUNCH=1): wait for every launch, so that no two kernels ever overlap
        static const bool sync_every = getenv("GADJ_SYNC_EVERY_LAUNCH") != nullptr;
        if (sync_every)
            dev::sync();
    }
}

void run_launches(gadj_ctx* c, const std::vector<Launch>& list)
{
    for (const Launch& L : list)
        run_one(c, L);
}

enum { PK_ASSEMBLE = 100, PK_OTHER = 101 };

// FormConstraintStationVarianceMatrix (ADJ:2041-2137): inverse-variance block, row-major
bool constraint_block(const gadj_ctx* c, const dna_stn_t& s, double* out)
{
    double varC = c->o.fixed_std_dev * c->o.fixed_std_dev;
    double varF = c->o.free_std_dev * c->o.free_std_dev;
    const char* k = s.stationConst;
    std::memset(out, 0, 9 * sizeof(double));
    if (k[0] == 'C' && k[1] == 'C' && k[2] == 'C') {
        out[0] = out[4] = out[8] = 1. / varC;
        return true;
    }
    if (k[0] == 'F' && k[1] == 'F' && k[2] == 'F') {
        out[0] = out[4] = out[8] = 1. / varF;
        return true;
    }
    double L[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    bool llh = (s.suppliedStationType == DNA_LLH_TYPE || s.suppliedStationType == DNA_LLh_TYPE);
    double v0 = (k[0] == 'F') ? varF : varC;
    double v1 = (k[1] == 'F') ? varF : varC;
    if (llh) {
        L[4] = v0;  // latitude  -> north
        L[0] = v1;  // longitude -> east
    } else {
        L[0] = v0;
        L[4] = v1;
    }
    L[8] = (k[2] == 'F') ? varF : varC;
    double V[9];
    if (s.suppliedStationType == DNA_XYZ_TYPE)
        std::memcpy(V, L, sizeof(V));
    else {
        double R[9], T[9];
        local_to_cart_rotation(s.currentLatitude, s.currentLongitude, R);
        mat3_mul(R, false, L, false, T);
        mat3_mul(T, false, R, true, V);
    }
    double up[6] = {V[0], V[1], V[2], V[4], V[5], V[8]}, inv[6];
    if (!spd3_inverse(up, inv))
        return false;
    out[0] = inv[0];
    out[1] = out[3] = inv[1];
    out[2] = out[6] = inv[2];
    out[4] = inv[3];
    out[5] = out[7] = inv[4];
    out[8] = inv[5];
    return true;
}

// LoadVarianceScaling + the scaling half of LoadVarianceMatrix_G (ADJ:4214-4282, 4453-4491):
// applied once, on the host, when the records are raw; the scaled variances are written back
// into the records so that every later consumer (device assembly, chi-square) reads them as-is.
void first_run_reduction(gadj_ctx* c)
{
    const double lim = std::fmin(1.0e-5, c->o.fixed_std_dev);
    dna_msr_t* msr = c->msr;
    const dna_stn_t* stn = c->stn;
    const Ellipsoid ell = c->ell;
    const std::vector<uint32_t>& first = c->first;
    const int reduced = c->reduced;
    parallel_for(first.size(), [&](uint64_t b0, uint64_t b1) {
        for (uint64_t b = b0; b < b1; ++b) {
            dna_msr_t* m = msr + first[b];
            if (reduced) {
                for (int r = 0; r < 3; ++r)
                    m[r].term1 = m[r].preAdjMeas;  // InitialiseMeasurement (ADJ:3913-3935)
                continue;
            }
            for (int r = 0; r < 3; ++r)
                m[r].preAdjMeas = m[r].term1;
            double vS = m[0].scale4, pS = m[0].scale1, lS = m[0].scale2, hS = m[0].scale3;
            if (vS < lim)
                vS = 1.0;
            if (pS < lim)
                pS = 1.0;
            if (lS < lim)
                lS = 1.0;
            if (hS < lim)
                hS = 1.0;
            bool scaleMatrix = std::fabs(vS - 1.0) > 1.0e-5;
            bool scalePartial =
                std::fabs(pS - 1.0) > 1.0e-5 || std::fabs(lS - 1.0) > 1.0e-5 || std::fabs(hS - 1.0) > 1.0e-5;
            if (!scaleMatrix && !scalePartial)
                continue;
            if (scalePartial && scaleMatrix) {
                pS *= vS;
                lS *= vS;
                hS *= vS;
            }
            double V[9];
            double s = scaleMatrix ? vS : 1.0;
            V[0] = m[0].term2 * s;
            V[1] = V[3] = m[1].term2 * s;
            V[4] = m[1].term3 * s;
            V[2] = V[6] = m[2].term2 * s;
            V[5] = V[7] = m[2].term3 * s;
            V[8] = m[2].term4 * s;
            if (scalePartial) {
                // ScaleGPSVCV (dnatemplatematrixfuncs.hpp:372-399) at station 1
                const dna_stn_t& s1 = stn[m[2].station1];
                double R[9], Ri[9], T[9], Vg[9];
                cart_geo_jacobian(ell, s1.currentLatitude, s1.currentLongitude, s1.currentHeight, R);
                if (mat3_inverse(R, Ri)) {
                    mat3_mul(Ri, false, V, false, T);
                    mat3_mul(T, false, Ri, true, Vg);
                    double sc[3] = {std::sqrt(pS), std::sqrt(lS), std::sqrt(hS)};
                    for (int i = 0; i < 3; ++i)
                        for (int j = 0; j < 3; ++j)
                            Vg[i * 3 + j] *= sc[i] * sc[j];
                    mat3_mul(R, false, Vg, false, T);
                    mat3_mul(T, false, R, true, V);
                }
            }
            m[0].term2 = V[0];
            m[1].term2 = V[1];
            m[1].term3 = V[4];
            m[2].term2 = V[2];
            m[2].term3 = V[5];
            m[2].term4 = V[8];
        }
    });
}

// GNSS point clusters 'Y' given as latitude / longitude / height become Cartesian on their first run
// (UpdateDesignNormalMeasMatrices_Y ADJ:6281-6325, LoadVarianceMatrix_Y ADJ:4563-4644): the original values are kept
// in preAdjMeas, orthometric heights are reduced with the station's geoid separation (kept in preAdjCorr), the
// records are rewritten in place as XYZ (station3 remembers the original type) and the cluster variance matrix is
// propagated with the Jacobians d(XYZ)/d(lat, lon, h) at the stations' current positions.  Written back into the
// caller's records, as the reference does.
int convert_llh_point_clusters(gadj_ctx* c)
{
    for (uint64_t i = 0; i < c->nmsr;) {
        dna_msr_t* m = &c->msr[i];
        if (m->measType != 'Y' || m->measStart != 0) {
            ++i;
            continue;
        }
        // members of this cluster
        const uint32_t count = m->vectorCount1;
        std::vector<uint64_t> rec;
        uint64_t j = i;
        for (uint32_t k = 0; k < count && j + 2 < c->nmsr; ++k) {
            rec.push_back(j);
            j += 3 + 3ull * c->msr[j].vectorCount2;
        }
        i = std::max<uint64_t>(j, i + 1);
        const bool LLH = std::strncmp(m->coordType, "LLH", 3) == 0, LLh = std::strncmp(m->coordType, "LLh", 3) == 0;
        if ((!LLH && !LLh) || m->ignore)
            continue;
        if (rec.size() != count || j > c->nmsr)
            return c->fail("truncated GNSS point cluster 'Y' (record " + std::to_string(rec.empty() ? j : rec[0]) + "): " +
                           std::to_string(count) + " points announced, the measurement list ends after " +
                           std::to_string(rec.size()));
        const uint32_t n = 3 * count;
        std::vector<double> V((size_t)n * n, 0.0), J((size_t)n * 3, 0.0);   // J: one 3x3 block per member
        auto sym = [&](uint32_t r, uint32_t col, double v) { V[(size_t)r * n + col] = V[(size_t)col * n + r] = v; };
        for (uint32_t k = 0; k < count; ++k) {
            dna_msr_t* r = &c->msr[rec[k]];
            if (r->station1 >= c->nstn)
                return c->fail("measurement refers to a station index beyond the station list");
            const dna_stn_t& st = c->stn[r->station1];
            const uint32_t v = 3 * k;
            sym(v, v, r[0].term2);
            sym(v, v + 1, r[1].term2);
            sym(v + 1, v + 1, r[1].term3);
            sym(v, v + 2, r[2].term2);
            sym(v + 1, v + 2, r[2].term3);
            sym(v + 2, v + 2, r[2].term4);
            for (uint32_t q = 0; q < r[0].vectorCount2 && v + 3 + 3 * q + 2 < n; ++q) {
                const dna_msr_t* cv = r + 3 + 3 * q;
                const uint32_t cc = v + 3 + 3 * q;
                for (uint32_t x = 0; x < 3; ++x) {
                    sym(v + x, cc, cv[x].term1);
                    sym(v + x, cc + 1, cv[x].term2);
                    sym(v + x, cc + 2, cv[x].term3);
                }
            }
            cart_geo_jacobian(c->ell, st.currentLatitude, st.currentLongitude, st.currentHeight, &J[9 * (size_t)k]);
            double h = r[2].term1;
            for (int q = 0; q < 3; ++q)
                r[q].preAdjMeas = r[q].term1;
            if (LLH && std::fabs((double)st.geoidSep) > 1.0e-4) {
                r[2].preAdjCorr = st.geoidSep;
                h += r[2].preAdjCorr;
            }
            double xyz[3];
            geo_to_cart(c->ell, r[0].term1, r[1].term1, h, xyz);
            for (int q = 0; q < 3; ++q) {
                r[q].term1 = xyz[q];
                std::snprintf(r[q].coordType, sizeof(r[q].coordType), "%s", "XYZ");
                r[q].station3 = LLH ? DNA_LLH_TYPE : DNA_LLh_TYPE;
            }
        }
        // block (a, b) of J V J^T = J_a V_ab J_b^T
        auto block = [&](uint32_t a, uint32_t b2, double* out) {
            double T[9];
            for (int x = 0; x < 3; ++x)
                for (int y = 0; y < 3; ++y) {
                    double sum = 0.0;
                    for (int z = 0; z < 3; ++z)
                        sum += J[9 * (size_t)a + 3 * x + z] * V[(size_t)(3 * a + z) * n + 3 * b2 + y];
                    T[3 * x + y] = sum;
                }
            for (int x = 0; x < 3; ++x)
                for (int y = 0; y < 3; ++y) {
                    double sum = 0.0;
                    for (int z = 0; z < 3; ++z)
                        sum += T[3 * x + z] * J[9 * (size_t)b2 + 3 * y + z];
                    out[3 * x + y] = sum;
                }
        };
        for (uint32_t k = 0; k < count; ++k) {
            dna_msr_t* r = &c->msr[rec[k]];
            double B[9];
            block(k, k, B);
            r[0].term2 = B[0];
            r[1].term2 = B[1];
            r[1].term3 = B[4];
            r[2].term2 = B[2];
            r[2].term3 = B[5];
            r[2].term4 = B[8];
            for (uint32_t q = 0; q < r[0].vectorCount2 && k + 1 + q < count; ++q) {
                dna_msr_t* cv = r + 3 + 3 * q;
                block(k, k + 1 + q, B);
                for (int x = 0; x < 3; ++x) {
                    cv[x].term1 = B[3 * x];
                    cv[x].term2 = B[3 * x + 1];
                    cv[x].term3 = B[3 * x + 2];
                }
            }
        }
    }
    return 0;
}

bool is_scalar_type(char t)
{
    switch (t) {
    case 'A': case 'B': case 'C': case 'E': case 'H': case 'I': case 'J': case 'K':
    case 'L': case 'M': case 'P': case 'Q': case 'R': case 'S': case 'V': case 'Z':
        return true;
    default:
        return false;
    }
}
int stations_of_type(char t)
{
    switch (t) {
    case 'A': case 'D':
        return 3;
    case 'H': case 'I': case 'J': case 'P': case 'Q': case 'R': case 'Y':
        return 1;
    default:
        return 2;
    }
}

// Scan the record list (the reference's CML: first record of each non-ignored measurement, ADJH:1216-1218) into
//   first[]      GNSS baselines 'G'
//   rows[]       one RowDesc per design row of every other type
//   clusters[]   D sets, X and Y clusters with their local station lists and station -> row incidence lists
int scan_measurements(gadj_ctx* c)
{
    c->first.clear();
    c->rows.clear();
    c->row_base.clear();
    c->clusters.clear();
    c->cstn.clear();
    c->inc_ptr.assign(1, 0u);
    c->inc.clear();
    c->non_gps = false;
    // a list of nothing but GNSS baselines (three records each) — the national-scale case — is scanned by all host threads
    if (c->nmsr % 3 == 0 && c->nmsr >= 3 * 65536) {
        const uint64_t nb = c->nmsr / 3;
        std::atomic<bool> pure{true};
        const unsigned nt = host_threads();
        std::vector<uint64_t> cnt(nt + 1, 0);
        const uint64_t chunk = (nb + nt - 1) / nt;
        parallel_for(nt, [&](uint64_t t0, uint64_t t1) {
            for (uint64_t t = t0; t < t1; ++t) {
                uint64_t n = 0;
                for (uint64_t b = t * chunk; b < std::min(nb, (t + 1) * chunk); ++b) {
                    const dna_msr_t* m = c->msr + 3 * b;
                    if (m[0].measType != 'G' || m[1].measType != 'G' || m[2].measType != 'G' || m[0].measStart != 0) {
                        pure.store(false, std::memory_order_relaxed);
                        return;
                    }
                    n += !m[0].ignore;
                }
                cnt[t + 1] = n;
            }
        }, 0);
        if (pure.load() && 3 * nb <= 0xFFFFFFF0ull) {
            for (unsigned t = 0; t < nt; ++t)
                cnt[t + 1] += cnt[t];
            c->first.resize(cnt[nt]);
            parallel_for(nt, [&](uint64_t t0, uint64_t t1) {
                for (uint64_t t = t0; t < t1; ++t) {
                    uint64_t o = cnt[t];
                    for (uint64_t b = t * chunk; b < std::min(nb, (t + 1) * chunk); ++b)
                        if (!c->msr[3 * b].ignore)
                            c->first[o++] = (uint32_t)(3 * b);
                }
            }, 0);
            c->nbsl = c->first.size();
            c->nrows = 0;
            if (c->nbsl == 0)
                return c->fail("No valid measurements to process. All measurements may be ignored.");   // LoadNetworkFiles (ADJ:10176)
            // first[] rises in steps of at least three records: no gaps exactly when the last one is where a gapless list ends
            c->contiguous = c->first[c->nbsl - 1] == c->first[0] + 3 * (c->nbsl - 1);
            return 0;
        }
    }
    uint64_t i = 0;
    auto bad_station = [&](uint32_t s) { return s >= c->nstn; };
    // close a cluster: local stations + incidence lists from rows [row0, rows.size())
    auto close_cluster = [&](char type, uint32_t row0) {
        ClusterDesc cd;
        std::memset(&cd, 0, sizeof(cd));
        cd.type = (uint8_t)type;
        cd.row0 = row0;
        cd.n = (uint32_t)c->rows.size() - row0;
        cd.st0 = (uint32_t)c->cstn.size();
        std::vector<uint32_t> loc;
        for (uint32_t r = row0; r < c->rows.size(); ++r)
            for (int k = 0; k < c->rows[r].nst; ++k)
                loc.push_back(c->rows[r].st[k]);
        std::sort(loc.begin(), loc.end());
        loc.erase(std::unique(loc.begin(), loc.end()), loc.end());
        cd.ns = (uint32_t)loc.size();
        std::vector<std::vector<uint32_t>> lists(loc.size());
        for (uint32_t r = row0; r < c->rows.size(); ++r)
            for (int k = 0; k < c->rows[r].nst; ++k) {
                uint32_t j = (uint32_t)(std::lower_bound(loc.begin(), loc.end(), c->rows[r].st[k]) - loc.begin());
                lists[j].push_back(((r - row0) << 2) | (uint32_t)k);
            }
        for (uint32_t j = 0; j < loc.size(); ++j) {
            c->cstn.push_back(loc[j]);
            c->inc.insert(c->inc.end(), lists[j].begin(), lists[j].end());
            c->inc_ptr.push_back((uint32_t)c->inc.size());
        }
        c->clusters.push_back(cd);
    };
    while (i < c->nmsr) {
        const dna_msr_t& m = c->msr[i];
        uint64_t step = 1;
        switch (m.measType) {
        case 'G':
            step = 3;
            break;
        case 'X':
        case 'Y': {
            uint64_t j = i;
            for (uint32_t k = 0; k < m.vectorCount1 && j < c->nmsr; ++k)
                j += 3 + 3ull * c->msr[j].vectorCount2;
            step = j - i;
            break;
        }
        case 'D':
            step = m.vectorCount1;   // RO record + target directions (dnadirectionset.cpp:430-466)
            break;
        default:
            step = 1;
        }
        if (step == 0)
            step = 1;
        if (i + step > c->nmsr)
            return c->fail("truncated measurement at the end of the measurement list");
        if (!m.ignore) {
            if (i + step > 0xFFFFFFF0ull)
                return c->fail("measurement index exceeds 32 bits");
            if (m.measType == 'G') {
                c->first.push_back((uint32_t)i);
            } else if (is_scalar_type(m.measType)) {
                RowDesc r;
                std::memset(&r, 0, sizeof(r));
                r.rec = (uint32_t)i;
                r.type = (uint8_t)m.measType;
                r.nst = (uint8_t)stations_of_type(m.measType);
                r.st[0] = m.station1;
                r.st[1] = r.nst > 1 ? m.station2 : m.station1;
                r.st[2] = r.nst > 2 ? m.station3 : m.station1;
                for (int k = 0; k < r.nst; ++k)
                    if (bad_station(r.st[k]))
                        return c->fail("measurement refers to a station index beyond the station list");
                if ((r.nst > 1 && r.st[0] == r.st[1]) || (r.nst > 2 && (r.st[0] == r.st[2] || r.st[1] == r.st[2])))
                    return c->fail(std::string("measurement of type '") + m.measType + "' with repeated stations");
                if (!(m.term2 > 0.0))
                    return c->fail("Invalid variance matrix: non-positive measurement variance");
                c->rows.push_back(r);
                c->row_base.push_back((uint32_t)i);
                c->non_gps = true;   // v_msrTally_.ContainsNonGPS() (ADJ:2457)
            } else if (m.measType == 'D') {
                // derived angles between consecutive non-ignored directions (ADJ:5088-5129).  vectorCount2 = the number of
                // directions that take part (dnaimport sets it, the RO included): a set that is not ignored needs two
                if (m.vectorCount2 < 2 || m.vectorCount2 > m.vectorCount1)
                    return c->fail("direction set (record " + std::to_string(i) + ") takes part in the adjustment with " +
                                   std::to_string(m.vectorCount2) + " of " + std::to_string(m.vectorCount1) +
                                   " directions: at least two non-ignored directions are needed (ignore the set otherwise)");
                const uint32_t row0 = (uint32_t)c->rows.size();
                uint64_t prev = i;
                for (uint64_t j = i + 1; j < i + step && c->rows.size() - row0 + 1 < m.vectorCount2; ++j) {
                    if (c->msr[j].ignore)
                        continue;
                    RowDesc r;
                    std::memset(&r, 0, sizeof(r));
                    r.rec = (uint32_t)j;
                    r.type = 'D';
                    r.nst = 3;
                    r.clustered = 1;
                    r.st[0] = c->msr[prev].station1;
                    r.st[1] = c->msr[prev].station2;
                    r.st[2] = c->msr[j].station2;
                    for (int k = 0; k < 3; ++k)
                        if (bad_station(r.st[k]))
                            return c->fail("measurement refers to a station index beyond the station list");
                    if (r.st[0] == r.st[1] || r.st[0] == r.st[2] || r.st[1] == r.st[2])
                        return c->fail("direction set with repeated stations in one derived angle");
                    if (!(c->msr[j].term2 > 0.0) || !(c->msr[prev].term2 > 0.0))
                        return c->fail("Invalid variance matrix: non-positive direction variance");
                    c->rows.push_back(r);
                    c->row_base.push_back((uint32_t)prev);
                    prev = j;
                }
                if (c->rows.size() > row0) {
                    close_cluster('D', row0);
                    c->non_gps = true;
                }
            } else if (m.measType == 'X' || m.measType == 'Y') {
                if (m.measType == 'Y' && std::strncmp(m.coordType, "XYZ", 3) != 0)
                    return c->fail("GNSS point cluster 'Y': coordinates must be XYZ, LLH or LLh");
                const uint32_t row0 = (uint32_t)c->rows.size();
                uint64_t j = i;
                for (uint32_t k = 0; k < m.vectorCount1; ++k) {
                    const dna_msr_t& b = c->msr[j];
                    if (bad_station(b.station1) || (m.measType == 'X' && bad_station(b.station2)))
                        return c->fail("measurement refers to a station index beyond the station list");
                    if (m.measType == 'X' && b.station1 == b.station2)
                        return c->fail("GNSS baseline with identical end stations");
                    for (int q = 0; q < 3; ++q) {
                        RowDesc r;
                        std::memset(&r, 0, sizeof(r));
                        r.rec = (uint32_t)(j + q);
                        r.type = (uint8_t)m.measType;
                        r.nst = m.measType == 'X' ? 2 : 1;
                        r.comp = (uint8_t)q;
                        r.clustered = 1;
                        r.st[0] = b.station1;
                        r.st[1] = m.measType == 'X' ? b.station2 : b.station1;
                        r.st[2] = b.station1;
                        c->rows.push_back(r);
                        c->row_base.push_back((uint32_t)j);
                    }
                    j += 3 + 3ull * b.vectorCount2;
                }
                if (c->rows.size() > row0)
                    close_cluster(m.measType, row0);
            } else
                return c->fail(std::string("measurement type '") + m.measType + "' is not a DynAdjust measurement type");
        }
        i += step;
    }
    c->nbsl = c->first.size();
    c->nrows = c->rows.size();
    if (c->nbsl + c->nrows == 0)
        return c->fail("No valid measurements to process. All measurements may be ignored.");   // LoadNetworkFiles (ADJ:10176)
    c->contiguous = true;
    for (uint64_t b = 0; b < c->nbsl; ++b)
        if (c->first[b] != c->first[0] + 3 * b) {
            c->contiguous = false;
            break;
        }
    return 0;
}

// First-run handling of the rows (host, once): back up / restore the raw value (InitialiseMeasurement, ADJ:3913-3935),
// deflection / geoid reductions per type (rows.h first_run_reduce), derived angles and their tridiagonal variance
// matrix for D sets (ADJ:5082-5193, ADJ:4059-4188), cluster variance matrices of X / Y with the whole-matrix scalar
// (ADJ:4312-4450, ADJ:4494-4679).  est: a-priori Cartesian coordinates.  Fills c->cvmat (V, to be inverted on the device).
int first_run_reduction_rows(gadj_ctx* c, const double* est)
{
    dna_msr_t* msr = c->msr;
    for (uint64_t r = 0; r < c->nrows; ++r) {
        const RowDesc& d = c->rows[r];
        if (d.clustered)
            continue;
        dna_msr_t& m = msr[d.rec];
        if (c->reduced)
            m.term1 = m.preAdjMeas;
        else
            m.preAdjMeas = m.term1;
        first_run_reduce((char)d.type, m, d.st, c->stn, est);
    }
    uint64_t pool = 0;
    for (ClusterDesc& cd : c->clusters) {
        cd.vinv_off = pool;
        pool += (uint64_t)cd.n * cd.n;
    }
    c->cvmat.assign(pool, 0.0);
    const double lim = std::fmin(1.0e-5, c->o.fixed_std_dev);
    for (ClusterDesc& cd : c->clusters) {
        double* V = c->cvmat.data() + cd.vinv_off;
        const uint32_t n = cd.n;
        if (cd.type == 'D') {
            const uint32_t ro = c->row_base[cd.row0];
            double previousDirection = msr[ro].term1;
            for (uint32_t a = 0; a < n; ++a) {
                const RowDesc& d = c->rows[cd.row0 + a];
                dna_msr_t& dir = msr[d.rec];
                dna_msr_t angle = msr[c->row_base[cd.row0 + a]];   // the reference's scratch angle record
                if (c->reduced)
                    angle.term1 = dir.preAdjMeas;
                else {
                    angle.term1 = dir.term1 - previousDirection;
                    if (angle.term1 < 0)
                        angle.term1 += kTwoPi;
                    if (angle.term1 > kTwoPi)
                        angle.term1 -= kTwoPi;
                }
                angle.preAdjMeas = angle.term1;
                first_run_reduce('D', angle, d.st, c->stn, est);
                dir.scale1 = angle.term1;
                dir.preAdjMeas = angle.preAdjMeas;
                dir.preAdjCorr = angle.preAdjCorr;
                previousDirection = dir.term1;
            }
            if (!c->reduced) {
                // angle = direction_a - direction_(a-1): V = A diag(var) A^T is tridiagonal
                double prevVar = msr[ro].term2;
                for (uint32_t a = 0; a < n; ++a) {
                    dna_msr_t& dir = msr[c->rows[cd.row0 + a].rec];
                    dir.scale2 = prevVar + dir.term2;
                    dir.scale3 = (a + 1 < n) ? -dir.term2 : 0.0;
                    prevVar = dir.term2;
                }
            }
            for (uint32_t a = 0; a < n; ++a) {
                const dna_msr_t& dir = msr[c->rows[cd.row0 + a].rec];
                V[(size_t)a * n + a] = dir.scale2;
                if (a + 1 < n)
                    V[(size_t)a * n + a + 1] = V[(size_t)(a + 1) * n + a] = dir.scale3;
            }
            continue;
        }
        // X / Y
        const uint32_t members = n / 3;
        dna_msr_t* m0 = &msr[c->rows[cd.row0].rec];
        double vS = m0->scale4, pS = m0->scale1, lS = m0->scale2, hS = m0->scale3;
        if (vS < lim)
            vS = 1.0;
        if (pS < lim)
            pS = 1.0;
        if (lS < lim)
            lS = 1.0;
        if (hS < lim)
            hS = 1.0;
        bool scaleMatrix = std::fabs(vS - 1.0) > 1.0e-5;
        bool scalePartial = std::fabs(pS - 1.0) > 1.0e-5 || std::fabs(lS - 1.0) > 1.0e-5 || std::fabs(hS - 1.0) > 1.0e-5;
        if (c->reduced)
            scaleMatrix = scalePartial = false;    // the records already hold the scaled matrix
        if (scalePartial && scaleMatrix) {         // LoadVarianceScaling (ADJ:4484-4490)
            pS *= vS;
            lS *= vS;
            hS *= vS;
        }
        // X clusters take the whole-matrix scalar while loading (ADJ:4358-4392) — also when partial scalars follow —
        // Y clusters afterwards and only without partial scalars (ADJ:4646-4647)
        const double onload = (cd.type == 'X' && scaleMatrix) ? vS : 1.0;
        auto put = [&](uint32_t r, uint32_t col, double field) { V[(size_t)r * n + col] = V[(size_t)col * n + r] = field * onload; };
        for (uint32_t k = 0; k < members; ++k) {
            dna_msr_t* r = &msr[c->rows[cd.row0 + 3 * k].rec];
            for (int q = 0; q < 3; ++q) {
                // clusters that arrived as latitude / longitude / height keep the original values in preAdjMeas (ADJ:6353)
                if (cd.type == 'Y' && (r[q].station3 == DNA_LLH_TYPE || r[q].station3 == DNA_LLh_TYPE))
                    continue;
                if (c->reduced)
                    r[q].term1 = r[q].preAdjMeas;
                else
                    r[q].preAdjMeas = r[q].term1;
            }
            const uint32_t v = 3 * k;
            put(v, v, r[0].term2);
            put(v, v + 1, r[1].term2);
            put(v + 1, v + 1, r[1].term3);
            put(v, v + 2, r[2].term2);
            put(v + 1, v + 2, r[2].term3);
            put(v + 2, v + 2, r[2].term4);
            const uint32_t ncov = r[0].vectorCount2;
            if (v + 3 + 3 * ncov > n)
                return c->fail("GNSS cluster covariance records exceed the cluster size");
            for (uint32_t q = 0; q < ncov; ++q) {
                dna_msr_t* cv = r + 3 + 3 * q;
                const uint32_t cc = v + 3 + 3 * q;
                for (uint32_t x = 0; x < 3; ++x) {
                    put(v + x, cc, cv[x].term1);
                    put(v + x, cc + 1, cv[x].term2);
                    put(v + x, cc + 2, cv[x].term3);
                }
            }
        }
        if (scalePartial) {
            // ScaleGPSVCV_Cluster (MFN:401-438): V' = (J S J^-1) V (J S J^-1)^T block by block, J = d(XYZ)/d(lat, lon, h) at
            // the member's first station, S = diag(sqrt p, sqrt l, sqrt h)
            std::vector<double> M((size_t)members * 9), W((size_t)n * n);
            for (uint32_t k = 0; k < members; ++k) {
                const dna_stn_t& st = c->stn[msr[c->rows[cd.row0 + 3 * k].rec].station1];
                double J[9], Ji[9];
                cart_geo_jacobian(c->ell, st.currentLatitude, st.currentLongitude, st.currentHeight, J);
                if (!mat3_inverse(J, Ji))
                    return c->fail("variance scaling: singular geographic Jacobian");
                const double sc[3] = {std::sqrt(pS), std::sqrt(lS), std::sqrt(hS)};
                for (int a = 0; a < 3; ++a)
                    for (int b2 = 0; b2 < 3; ++b2)
                        M[9 * (size_t)k + 3 * a + b2] = J[3 * a] * sc[0] * Ji[b2] + J[3 * a + 1] * sc[1] * Ji[3 + b2] + J[3 * a + 2] * sc[2] * Ji[6 + b2];
            }
            for (uint32_t ka = 0; ka < members; ++ka)
                for (uint32_t kb = 0; kb < members; ++kb) {
                    double T[9];
                    for (int x = 0; x < 3; ++x)
                        for (int y = 0; y < 3; ++y)
                            T[3 * x + y] = M[9 * (size_t)ka + 3 * x] * V[(size_t)(3 * ka) * n + 3 * kb + y] +
                                           M[9 * (size_t)ka + 3 * x + 1] * V[(size_t)(3 * ka + 1) * n + 3 * kb + y] +
                                           M[9 * (size_t)ka + 3 * x + 2] * V[(size_t)(3 * ka + 2) * n + 3 * kb + y];
                    for (int x = 0; x < 3; ++x)
                        for (int y = 0; y < 3; ++y)
                            W[(size_t)(3 * ka + x) * n + 3 * kb + y] = T[3 * x] * M[9 * (size_t)kb + 3 * y] + T[3 * x + 1] * M[9 * (size_t)kb + 3 * y + 1] +
                                                                   T[3 * x + 2] * M[9 * (size_t)kb + 3 * y + 2];
                }
            std::copy(W.begin(), W.end(), V);
        } else if (cd.type == 'Y' && scaleMatrix) {
            for (size_t x = 0; x < (size_t)n * n; ++x)
                V[x] *= vS;
        }
        if (scaleMatrix || scalePartial) {
            // written back: later consumers read the scaled matrix (SetGPSVarianceMatrix, ADJ:4425, 4654)
            for (uint32_t k = 0; k < members; ++k) {
                dna_msr_t* r = &msr[c->rows[cd.row0 + 3 * k].rec];
                const uint32_t v = 3 * k;
                r[0].term2 = V[(size_t)v * n + v];
                r[1].term2 = V[(size_t)v * n + v + 1];
                r[1].term3 = V[(size_t)(v + 1) * n + v + 1];
                r[2].term2 = V[(size_t)v * n + v + 2];
                r[2].term3 = V[(size_t)(v + 1) * n + v + 2];
                r[2].term4 = V[(size_t)(v + 2) * n + v + 2];
                for (uint32_t q = 0; q < r[0].vectorCount2; ++q) {
                    dna_msr_t* cv = r + 3 + 3 * q;
                    const uint32_t cc = v + 3 + 3 * q;
                    for (uint32_t x = 0; x < 3; ++x) {
                        cv[x].term1 = V[(size_t)(v + x) * n + cc];
                        cv[x].term2 = V[(size_t)(v + x) * n + cc + 1];
                        cv[x].term3 = V[(size_t)(v + x) * n + cc + 2];
                    }
                }
            }
        }
    }
    return 0;
}

}  // namespace

extern "C" {

void gadj_default_opts(gadj_opts* o)
{
    std::memset(o, 0, sizeof(*o));
    o->fixed_std_dev = 1.0e-6;
    o->free_std_dev = 10.0;
    o->iteration_threshold = (double)0.0005f;
    o->semi_major = 6378137.0;
    o->inv_flattening = 298.257222101;
    o->confidence_interval = 95.0;
    o->workspace_gb = 0.0;
    o->max_iterations = 10;
    o->scale_normals_to_unity = 1;
    o->ordering = GADJ_ORDER_AUTO;
    o->leaf_stations = 96;
    o->device = 0;
}

const char* gadj_last_error(const gadj_ctx* c) { return c ? c->err.c_str() : g_create_error.c_str(); }

int gadj_create(const gadj_opts* o, gadj_ctx** out)
{
    *out = nullptr;
    gadj_opts d;
    gadj_default_opts(&d);
    if (o)
        d = *o;
    std::string e;
    dev::Device* device = dev::open(d.device, e);
    if (!device) {
        g_create_error = e;
        return 1;
    }
    gadj_ctx* c = new gadj_ctx();
    c->device = device;
    c->o = d;
    if (c->o.leaf_stations == 0)
        c->o.leaf_stations = 96;
    c->ell = make_ellipsoid(c->o.semi_major, c->o.inv_flattening);
    double conf = c->o.confidence_interval * 0.01;
    conf += (1.0 - conf) / 2.0;
    c->critical = norm_quantile(conf);
    for (auto& e2 : c->ev)
        e2 = dev::event_create();
    *out = c;
    return 0;
}

void gadj_destroy(gadj_ctx* c)
{
    if (!c)
        return;
    dev::use(c->device);
    dev::sync();
    for (auto& e : c->ev)
        if (e)
            dev::event_destroy(e);
    for (auto& e : c->prof_ev)
        dev::event_destroy(e);
    for (auto& m : c->peer_maps)
        dev::peer_unmap(m.first, m.second);
    dev::Device* device = c->device;
    delete c;          // frees the device arrays while the device is still current
    dev::close(device);
}

int gadj_set_stations(gadj_ctx* c, dna_stn_t* stn, uint32_t count)
{
    dev::use(c->device);
    if (!stn || count == 0)
        return c->fail("empty station list");
    c->stn = stn;
    c->nstn = count;
    c->prepared = false;
    return 0;
}

int gadj_set_measurements(gadj_ctx* c, dna_msr_t* msr, uint64_t count)
{
    dev::use(c->device);
    if (!msr || count == 0)
        return c->fail("empty measurement list");
    c->msr = msr;
    c->nmsr = count;
    c->prepared = false;
    return 0;
}

int gadj_set_measurements_reduced(gadj_ctx* c, int reduced)
{
    dev::use(c->device);
    c->reduced = reduced ? 1 : 0;
    c->prepared = false;
    return 0;
}

int gadj_set_blocks(gadj_ctx* c, uint32_t nblocks, const uint32_t* isl_off, const uint32_t* isl)
{
    dev::use(c->device);
    c->isl_off.clear();
    c->isl.clear();
    if (nblocks > 1) {
        if (!isl_off || !isl)
            return c->fail("block lists missing");
        c->isl_off.assign(isl_off, isl_off + nblocks + 1);
        c->isl.assign(isl, isl + isl_off[nblocks]);
    }
    c->prepared = false;
    return 0;
}

int gadj_prepare(gadj_ctx* c)
{
    dev::use(c->device);
    // GADJ_DEBUG: wall time of the host phases
    const bool timing = getenv("GADJ_DEBUG") != nullptr;
    auto tick = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (!timing)
            return;
        auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "gadj_prepare[%d]: %-28s %8.1f ms\n", c->mg_rank, what, std::chrono::duration<double, std::milli>(now - tick).count());
        tick = now;
    };
    c->prepared = false;
    c->factor_valid = c->inverse_valid = c->normals_valid = false;
    c->iteration = 0;
    if (!c->stn || !c->msr)
        return c->fail("stations and measurements must be set before gadj_prepare");
    if (!c->reduced && convert_llh_point_clusters(c))
        return 1;
    if (scan_measurements(c))
        return 1;
    {
        std::atomic<int> bad_bsl{0};
        parallel_for(c->nbsl, [&](uint64_t b0, uint64_t b1) {
            for (uint64_t b = b0; b < b1; ++b) {
                const dna_msr_t& m = c->msr[c->first[b]];
                if (m.station1 >= c->nstn || m.station2 >= c->nstn)
                    bad_bsl.store(1, std::memory_order_relaxed);
                else if (m.station1 == m.station2)
                    bad_bsl.fetch_or(2, std::memory_order_relaxed);
            }
        });
        if (bad_bsl.load() & 1)
            return c->fail("measurement refers to a station index beyond the station list");
        if (bad_bsl.load() & 2)
            return c->fail("GNSS baseline with identical end stations");
    }
    // a-priori Cartesian coordinates (PopulateEstimatedStationMatrix, ADJ:632-693): the first-run reductions need them
    std::vector<double> est(3 * (size_t)c->nstn);
    parallel_for(c->nstn, [&](uint64_t s0, uint64_t s1) {
        for (uint64_t s = s0; s < s1; ++s)
            geo_to_cart(c->ell, c->stn[s].currentLatitude, c->stn[s].currentLongitude, c->stn[s].currentHeight, &est[3 * s]);
    });
    lap("scan + a-priori coordinates");
    first_run_reduction(c);
    if (first_run_reduction_rows(c, est.data()))
        return 1;
    lap("first-run reductions");

    // unique station pairs -> edge slots: baselines, the station pairs of every independent row, every pair of a cluster
    const uint64_t nb = c->nbsl;
    auto pair_key = [](uint32_t a, uint32_t b2) { return ((uint64_t)std::min(a, b2) << 32) | std::max(a, b2); };
    std::vector<uint64_t> keys(nb);
    std::vector<uint32_t> end1(nb), end2(nb);   // the baselines' end stations, once: the 208-byte records are not walked again
    parallel_for(nb, [&](uint64_t b0, uint64_t b1) {
        for (uint64_t b = b0; b < b1; ++b) {
            const dna_msr_t& m = c->msr[c->first[b]];
            end1[b] = m.station1;
            end2[b] = m.station2;
            keys[b] = pair_key(m.station1, m.station2);
        }
    });
    for (const RowDesc& d : c->rows) {
        if (d.clustered)
            continue;
        for (int x = 0; x < d.nst; ++x)
            for (int y = x + 1; y < d.nst; ++y)
                keys.push_back(pair_key(d.st[x], d.st[y]));
    }
    for (const ClusterDesc& cd : c->clusters)
        for (uint32_t x = 0; x < cd.ns; ++x)
            for (uint32_t y = x + 1; y < cd.ns; ++y)
                keys.push_back(pair_key(c->cstn[cd.st0 + x], c->cstn[cd.st0 + y]));
    // distinct pairs, and how many measurements contribute to each: a pair fed by a single GNSS baseline is stored, not added
    std::vector<uint64_t> uniq(keys);
    parallel_sort(uniq);
    std::vector<uint32_t> pair_count;
    {
        size_t o = 0;
        for (size_t i = 0; i < uniq.size();) {
            size_t j = i;
            while (j < uniq.size() && uniq[j] == uniq[i])
                ++j;
            uniq[o++] = uniq[i];
            pair_count.push_back((uint32_t)(j - i));
            i = j;
        }
        uniq.resize(o);
    }
    c->nedge = uniq.size();
    if (c->nedge >= (1ull << 30))
        return c->fail("too many distinct station pairs");
    std::vector<std::pair<uint32_t, uint32_t>> edges(c->nedge);
    for (uint64_t e = 0; e < c->nedge; ++e)
        edges[e] = {(uint32_t)(uniq[e] >> 32), (uint32_t)(uniq[e] & 0xFFFFFFFFu)};

    std::vector<double> lat(c->nstn), lon(c->nstn);
    // stations no measurement that takes part touches are not parameters of the adjustment: the reference leaves them out
    // of its station lists (RemoveInvalidStations LDR:286-300, unknownParams_ ADJ:632-693); here they stay in the system,
    // held by their a-priori weight alone, but do not count as unknowns
    std::vector<uint8_t> used(c->nstn, 0);
    for (uint64_t b = 0; b < nb; ++b)
        used[end1[b]] = used[end2[b]] = 1;
    for (const RowDesc& r : c->rows)
        for (int k = 0; k < r.nst; ++k)
            used[r.st[k]] = 1;
    c->constrained_components = 0;
    c->unused_stations = 0;
    for (uint32_t s = 0; s < c->nstn; ++s) {
        lat[s] = c->stn[s].currentLatitude;
        lon[s] = c->stn[s].currentLongitude;
        if (!used[s]) {
            c->unused_stations++;
            continue;
        }
        for (int k = 0; k < 3; ++k)
            c->constrained_components += c->stn[s].stationConst[k] == 'C';
    }
    lap("station pairs");
    OrderingOptions oo;
    oo.leaf_stations = c->o.leaf_stations;
    oo.dense = c->o.ordering == GADJ_ORDER_DENSE;
    uint32_t nblocks = c->isl_off.empty() ? 1u : (uint32_t)c->isl_off.size() - 1;
    std::string e = analyse(c->nstn, edges, lat.data(), lon.data(), oo, nblocks,
                            c->isl_off.empty() ? nullptr : c->isl_off.data(), c->isl.empty() ? nullptr : c->isl.data(),
                            c->sym);
    if (!e.empty())
        return c->fail(e);
    if (c->mg_world > 1)
        finalize_layout(c->sym, c->mg_world, c->mg_rank);
    lap("ordering + symbolic");
    const Symbolic& S = c->sym;
    if (const char* fp = getenv("GADJ_DUMP_FRONTS")) {   // ordering studies: level, own unknowns, boundary unknowns, flops
        if (FILE* f = fopen(fp, "w")) {
            for (const Front& fr : S.fronts)
                fprintf(f, "%d,%u,%u,%.6e,%llu,%u,%u\n", fr.level, fr.k, fr.r, fr.work, (unsigned long long)fr.panel_off, fr.ldk,
                        fr.own_begin);
            fclose(f);
        }
    }

    // per-edge orientation and destinations
    c->edge_hi.resize(c->nedge);
    c->edge_lo.resize(c->nedge);
    std::vector<uint64_t> off_dest(c->nedge), diag_dest(c->nstn);
    std::vector<uint32_t> off_ld(c->nedge), diag_ld(c->nstn);
    std::atomic<bool> bad{false};
    parallel_for(c->nedge, [&](uint64_t e0, uint64_t e1) {
        for (uint64_t ei = e0; ei < e1; ++ei) {
            uint32_t a = edges[ei].first, b2 = edges[ei].second;
            uint32_t pa = S.pos_of_stn[a], pb = S.pos_of_stn[b2];
            uint32_t hi = pa > pb ? a : b2, lo = pa > pb ? b2 : a;
            c->edge_hi[ei] = hi;
            c->edge_lo[ei] = lo;
            uint64_t slot = find_slot(S, std::max(pa, pb), std::min(pa, pb));
            if (slot == UINT64_MAX) {
                bad.store(true, std::memory_order_relaxed);
                continue;
            }
            off_dest[ei] = S.ndest[slot];
            off_ld[ei] = S.ndest_ld[slot];
        }
    });
    if (bad)
        return c->fail("internal: station pair missing from the normal-matrix pattern");
    for (uint32_t s = 0; s < c->nstn; ++s) {
        uint64_t slot = S.ncol_ptr[S.pos_of_stn[s]];
        diag_dest[s] = S.ndest[slot];
        diag_ld[s] = S.ndest_ld[slot];
    }
    // edge word of the ordered pair (a, b): slot | bit31 when a is eliminated after b (a owns the rows of the stored block)
    auto edge_word_of = [&](uint32_t a, uint32_t b2) {
        uint64_t ei = std::lower_bound(uniq.begin(), uniq.end(), pair_key(a, b2)) - uniq.begin();
        return (uint32_t)ei | (S.pos_of_stn[a] > S.pos_of_stn[b2] ? 0x80000000u : 0u);
    };
    c->edge_word.resize(nb);
    parallel_for(nb, [&](uint64_t b0, uint64_t b1) {
        for (uint64_t b = b0; b < b1; ++b) {
            uint32_t w = edge_word_of(end1[b], end2[b]);
            if (pair_count[w & EDGE_SLOT_MASK] == 1)
                w |= EDGE_EXCLUSIVE;
            c->edge_word[b] = w;
        }
    });
    c->edge_bsl.assign(c->nedge, ~0u);
    for (uint64_t b = 0; b < nb; ++b)
        if (c->edge_word[b] & EDGE_EXCLUSIVE)
            c->edge_bsl[c->edge_word[b] & EDGE_SLOT_MASK] = (uint32_t)b;
    // incidence lists of the stations over the GNSS baselines, ascending baseline index (fixed summation order)
    c->binc_ptr.assign((size_t)c->nstn + 1, 0);
    for (uint64_t b = 0; b < nb; ++b) {
        ++c->binc_ptr[end1[b] + 1];
        ++c->binc_ptr[end2[b] + 1];
    }
    for (uint32_t s2 = 0; s2 < c->nstn; ++s2)
        c->binc_ptr[s2 + 1] += c->binc_ptr[s2];
    c->binc.resize(2 * nb);
    {
        std::vector<uint32_t> cur(c->binc_ptr.begin(), c->binc_ptr.end() - 1);
        for (uint64_t b = 0; b < nb; ++b) {
            c->binc[cur[end1[b]]++] = (uint32_t)b;
            c->binc[cur[end2[b]]++] = (uint32_t)b | 0x80000000u;
        }
    }
    for (RowDesc& d : c->rows) {
        if (d.nst >= 2)
            d.edge[0] = edge_word_of(d.st[0], d.st[1]);
        if (d.nst >= 3) {
            d.edge[1] = edge_word_of(d.st[0], d.st[2]);
            d.edge[2] = edge_word_of(d.st[1], d.st[2]);
        }
    }
    c->pair_word.clear();
    for (ClusterDesc& cd : c->clusters) {
        cd.pair_off = c->pair_word.size();
        for (uint32_t x = 0; x < cd.ns; ++x)
            for (uint32_t y = 0; y < cd.ns; ++y)
                c->pair_word.push_back(x == y ? 0u : edge_word_of(c->cstn[cd.st0 + x], c->cstn[cd.st0 + y]));
    }

    // geographic coordinates, geoid separations and constraint blocks
    std::vector<double> cb(9 * (size_t)c->nstn), llh0(3 * (size_t)c->nstn);
    std::vector<float> geoid(c->nstn);
    for (uint32_t s = 0; s < c->nstn; ++s) {
        llh0[3 * s] = c->stn[s].currentLatitude;
        llh0[3 * s + 1] = c->stn[s].currentLongitude;
        llh0[3 * s + 2] = c->stn[s].currentHeight;
        geoid[s] = c->stn[s].geoidSep;
        if (!constraint_block(c, c->stn[s], &cb[9 * s]))
            return c->fail("station constraint variance matrix is not positive definite");
    }

    lap("edge words, incidence lists");
    // ---- device memory ---------------------------------------------------------
    build_rowidx(S, c->plan);
    std::vector<double> z;
    bool ok = true;
    const bool mg = c->mg_world > 1;   // buffers the other ranks read / write over NVLink are allocated peer-visible
    c->connected = false;
    ok &= c->d_msr.resize(c->nmsr + 64, mg);
    ok &= c->d_first.upload(c->first);
    ok &= c->d_edge.upload(c->edge_word);
    ok &= c->d_binc_ptr.upload(c->binc_ptr);
    ok &= c->d_binc.upload(c->binc);
    ok &= c->d_edge_bsl.upload(c->edge_bsl);
    ok &= c->d_bq.resize(9 * (size_t)nb);
    ok &= c->d_rows.upload(c->rows);
    ok &= c->d_clusters.upload(c->clusters);
    ok &= c->d_cstn.upload(c->cstn);
    ok &= c->d_inc_ptr.upload(c->inc_ptr);
    ok &= c->d_inc.upload(c->inc);
    ok &= c->d_pair_word.upload(c->pair_word);
    ok &= c->d_geoid.upload(geoid);
    ok &= c->d_row_l.resize(c->nrows);
    ok &= c->d_row_a.resize(9 * (size_t)c->nrows);
    ok &= c->d_row_t.resize(c->nrows);
    ok &= c->d_cvinv.resize(c->cvmat.size());
    ok &= c->d_edge_hi.upload(c->edge_hi);
    ok &= c->d_edge_lo.upload(c->edge_lo);
    ok &= c->d_pos.upload(S.pos_of_stn);
    ok &= c->d_pos_owned.upload(S.pos_owned);
    ok &= c->d_diag_dest.upload(diag_dest);
    ok &= c->d_diag_ld.upload(diag_ld);
    ok &= c->d_off_dest.upload(off_dest);
    ok &= c->d_off_ld.upload(off_ld);
    ok &= c->d_est.upload(est);
    ok &= c->d_est0.upload(est);
    ok &= c->d_cblock.upload(cb);
    ok &= c->d_llh.upload(llh0);
    ok &= c->d_llh0.upload(llh0);
    ok &= c->d_ndiag.resize(9 * (size_t)c->nstn);
    ok &= c->d_noff.resize(9 * (size_t)c->nedge);
    ok &= c->d_w.resize(3 * (size_t)c->nstn);
    ok &= c->d_dscale.resize(3 * (size_t)c->nstn);
    ok &= c->d_x.resize(3 * (size_t)c->nstn, mg);
    ok &= c->d_y.resize(3 * (size_t)c->nstn);
    ok &= c->d_wbuf.resize(wbuf_doubles(S), mg);
    ok &= c->d_corr.resize(3 * (size_t)c->nstn + 8);
    ok &= c->d_vcvd.resize(9 * (size_t)c->nstn, mg);
    ok &= c->d_vcvo.resize(9 * (size_t)c->nedge, mg);
    ok &= c->d_sums.resize(8);
    ok &= c->d_info.resize(4, mg);
    ok &= c->d_counter.resize(2, mg);
    ok &= c->d_apply.resize(APPLY_SCRATCH_DOUBLES);
    ok &= c->d_rowmap.upload(S.rowmap);
    ok &= c->d_tgt.resize(S.targets.size());
    ok &= c->d_coltgt.resize(S.bnd.size());
    ok &= c->d_rowidx.upload(c->plan.rowidx);
    ok &= c->d_panels.resize(S.panel_doubles, mg);
    if (!ok)
        return c->fail("out of device memory while allocating the adjustment state");
    dev::h2d(c->d_msr.p, c->msr, c->nmsr * sizeof(dna_msr_t));
    if (!c->clusters.empty()) {
        // FormInverseVarianceMatrix (ADJ:8472-8517) of every cluster matrix, on the device
        DevArray<double> work;
        if (!work.upload(c->cvmat))
            return c->fail("out of device memory while inverting the cluster variance matrices");
        dev::zero(c->d_info.p, c->d_info.bytes());
        launch_cluster_inverse(c->d_clusters.p, (uint32_t)c->clusters.size(), work.p, c->d_cvinv.p, c->d_info.p, dev::stream());
        int32_t info[4] = {0, 0, 0, 0};
        dev::d2h(info, c->d_info.p, sizeof(info));
        std::string e2 = dev::sync();
        if (!e2.empty())
            return c->fail(e2);
        if (info[0] != 0) {
            const ClusterDesc& cd = c->clusters[info[0] - 1];
            return c->fail(std::string("Invalid variance matrix: the variance matrix of a '") + (char)cd.type +
                           "' cluster is not positive definite (record " + std::to_string(c->rows[cd.row0].rec) + ")");
        }
        c->cvmat.clear();
        c->cvmat.shrink_to_fit();
    }

    lap("device arrays + upload");
    size_t want = ideal_pool_doubles(S), least = min_pool_doubles(S);
    size_t budget;
    if (c->o.workspace_gb > 0)
        budget = (size_t)(c->o.workspace_gb * 1e9 / 8);
    else
        budget = (size_t)(0.6 * (double)dev::mem_free() / 8);
    size_t pool = std::max(least, std::min(want, budget));
    if (!c->d_pool.resize(pool, mg))
        return c->fail("out of device memory while allocating the inverse workspace");

    PlanBuffers pb;
    pb.panels = c->d_panels.p;
    pb.pool = c->d_pool.p;
    pb.pool_doubles = pool;
    pb.x = c->d_x.p;
    pb.y = c->d_y.p;
    pb.wbuf = c->d_wbuf.p;
    pb.rowmap = c->d_rowmap.p;
    pb.rowidx = c->d_rowidx.p;
    pb.tgt = c->d_tgt.p;
    pb.coltgt = c->d_coltgt.p;
    pb.gemm_tile = c->o.gemm_tile;
    if (const char* gt = getenv("GADJ_GEMM_TILE"))   // tuning aid: overrides the option
        pb.gemm_tile = atoi(gt);
    if (pb.gemm_tile != 0 && pb.gemm_tile != 64 && pb.gemm_tile != 128)
        return c->fail("gemm_tile must be 0 (automatic), 64 or 128");
    e = build_plan(S, pb, c->plan);
    if (!e.empty())
        return c->fail(e);
    lap("launch plan");
    ok = true;
    ok &= c->d_gemm.upload(c->plan.gemm);
    ok &= c->d_tiles.upload(c->plan.tiles);
    // the scatter tables were allocated before build_plan (the ops point into them): fill in place
    if (!c->plan.tgt.empty())
        dev::h2d(c->d_tgt.p, c->plan.tgt.data(), c->plan.tgt.size() * sizeof(ScatterTarget));
    if (!c->plan.coltgt.empty())
        dev::h2d(c->d_coltgt.p, c->plan.coltgt.data(), c->plan.coltgt.size() * sizeof(int32_t));
    ok &= c->d_diag.upload(c->plan.diag);
    ok &= c->d_tri.upload(c->plan.tri);
    ok &= c->d_gemv.upload(c->plan.gemv);
    ok &= c->d_tr.upload(c->plan.transpose);
    ok &= c->d_gather.upload(c->plan.gather);
    ok &= c->d_gather_tiles.upload(c->plan.gather_tiles);
    ok &= c->d_reduce.upload(c->plan.reduce);
    ok &= c->d_push.upload(c->plan.push);
    {
        std::vector<double*> bases(MC_BUFS, nullptr);
        bases[MC_PANELS] = c->d_panels.p;
        bases[MC_WBUF] = c->d_wbuf.p;
        bases[MC_POOL] = c->d_pool.p;
        bases[MC_X] = c->d_x.p;
        bases[MC_VCVD] = c->d_vcvd.p;
        bases[MC_VCVO] = c->d_vcvo.p;
        ok &= c->d_bases.upload(bases);
    }
    {
        // whole-vector reductions of a multi-GPU run: the solution vector, the station / pair variance blocks
        std::vector<ReduceOp> misc = {ReduceOp{0, 3 * (uint64_t)c->nstn}, ReduceOp{0, 9 * (uint64_t)c->nstn},
                                      ReduceOp{0, 9 * (uint64_t)c->nedge}};
        ok &= c->d_reduce_misc.upload(misc);
    }
    dev::zero(c->d_info.p, c->d_info.bytes());
    // the barrier counter starts at zero before any peer can see it (the handles are exchanged after gadj_prepare)
    if (c->d_counter.p)
        dev::zero(c->d_counter.p, c->d_counter.bytes());
    c->barrier_round = 0;
    if (!ok)
        return c->fail("out of device memory while uploading the launch plan");
    e = dev::sync();
    if (!e.empty())
        return c->fail(e);
    c->device_bytes = c->d_msr.bytes() + c->d_panels.bytes() + c->d_pool.bytes() + c->d_wbuf.bytes() + c->d_noff.bytes() + c->d_ndiag.bytes() +
                      c->d_gemm.bytes() + c->d_gemv.bytes() + c->d_vcvo.bytes() + c->d_vcvd.bytes();
    c->h_corr.assign(3 * (size_t)c->nstn, 0.0);
    lap("plan upload");
    c->prepared = true;
    return 0;
}

int gadj_get_info(const gadj_ctx* c, gadj_info* info)
{
    dev::use(c->device);
    std::memset(info, 0, sizeof(*info));
    if (!c->prepared)
        return 1;
    info->nstations = c->nstn;
    info->nbaselines = c->nbsl;
    info->nedges = c->nedge;
    info->nfronts = c->sym.fronts.size();
    info->nlevels = c->sym.levels.size();
    info->panel_bytes = c->sym.panel_doubles * 8;
    info->pool_bytes = c->d_pool.bytes();
    info->device_bytes = c->device_bytes;
    info->factor_flops = c->sym.factor_flops;      // whole network, from the symbolic factorisation
    info->inverse_flops = c->sym.inverse_flops;
    info->rank_factor_flops = c->sym.world > 1 ? c->sym.my_factor_flops : c->sym.factor_flops;
    info->rank_inverse_flops = c->sym.world > 1 ? c->sym.my_inverse_flops : c->sym.inverse_flops;
    info->cut_level = c->sym.world > 1 ? c->sym.cut_level : -1;
    info->top_fronts = 0;
    info->nvlink_read_bytes = c->plan.nvlink_read_bytes;
    info->nvlink_write_bytes = c->plan.nvlink_write_bytes;
    info->barriers_per_step = c->plan.barriers;
    for (const Front& f : c->sym.fronts)
        info->top_fronts += f.top;
    info->launches_factor = c->plan.factor.size();
    info->launches_solve = c->plan.fwd.size() + c->plan.bwd.size();
    info->launches_inverse = c->plan.selinv.size();
    for (const Front& f : c->sym.fronts) {
        info->max_front_rows = std::max(info->max_front_rows, f.m);
        info->max_front_cols = std::max(info->max_front_cols, f.k);
    }
    return 0;
}

int gadj_upload_measurements(gadj_ctx* c)
{
    dev::use(c->device);
    if (!c->prepared)
        return c->fail("gadj_prepare has not been run");
    if (c->mg_world <= 1 || !c->connected) {
        dev::h2d(c->d_msr.p, c->msr, c->nmsr * sizeof(dna_msr_t));
        return 0;
    }
    // multi-GPU: every rank sends 1/world of the list over its own PCIe link; the other shares are pulled from the peers'
    // device copies over NVLink (every rank assembles from the whole list)
    const uint64_t chunk = (c->nmsr + c->mg_world - 1) / c->mg_world;
    auto range = [&](int q, uint64_t& first, uint64_t& count) {
        first = std::min<uint64_t>(c->nmsr, (uint64_t)q * chunk);
        count = std::min<uint64_t>(chunk, c->nmsr - first);
    };
    uint64_t first, count;
    range(c->mg_rank, first, count);
    mg_barrier(c);   // nobody still assembles from the previous copy
    if (count)
        dev::h2d(c->d_msr.p + first, c->msr + first, count * sizeof(dna_msr_t));
    mg_barrier(c);
    for (int q = 0; q < c->mg_world; ++q) {
        if (q == c->mg_rank)
            continue;
        range(q, first, count);
        if (count)
            dev::d2d(c->d_msr.p + first, (const char*)(c->d_msr.p + first) + c->peers.delta[MC_MSR][q], count * sizeof(dna_msr_t));
    }
    mg_barrier(c);
    return 0;
}

int gadj_upload_measurements_range(gadj_ctx* c, uint64_t first, uint64_t count)
{
    dev::use(c->device);
    if (!c->prepared)
        return c->fail("gadj_prepare has not been run");
    if (first > c->nmsr || count > c->nmsr - first)
        return c->fail("measurement record range out of bounds");
    if (count)
        dev::h2d(c->d_msr.p + first, c->msr + first, count * sizeof(dna_msr_t));
    return 0;
}

int gadj_reset_estimates(gadj_ctx* c)
{
    dev::use(c->device);
    if (!c->prepared)
        return c->fail("gadj_prepare has not been run");
    dev::d2d(c->d_est.p, c->d_est0.p, c->d_est.bytes());
    dev::d2d(c->d_llh.p, c->d_llh0.p, c->d_llh.bytes());
    c->iteration = 0;
    c->factor_valid = c->inverse_valid = false;
    return 0;
}

static void fill_assemble(gadj_ctx* c, AssembleParams& ap, int normals)
{
    ap.msr = c->d_msr.p;
    ap.first = c->d_first.p;
    ap.edge = c->d_edge.p;
    ap.est = c->d_est.p;
    ap.bq = c->d_bq.p;
    ap.inc_ptr = c->d_binc_ptr.p;
    ap.inc = c->d_binc.p;
    ap.ndiag = c->d_ndiag.p;
    ap.noff = c->d_noff.p;
    ap.w = c->d_w.p;
    ap.nbaselines = c->nbsl;
    ap.nstn = c->nstn;
    ap.contiguous = c->contiguous ? 1 : 0;
    ap.normals = normals;
}

static void fill_rows(gadj_ctx* c, RowsParams& rp, int normals, int assemble)
{
    rp.msr = c->d_msr.p;
    rp.rows = c->d_rows.p;
    rp.nrows = c->nrows;
    rp.est = c->d_est.p;
    rp.llh = c->d_llh.p;
    rp.geoid = c->d_geoid.p;
    rp.row_l = c->d_row_l.p;
    rp.row_a = c->d_row_a.p;
    rp.ndiag = c->d_ndiag.p;
    rp.noff = c->d_noff.p;
    rp.w = c->d_w.p;
    rp.vcv_diag = c->d_vcvd.p;
    rp.vcv_off = c->d_vcvo.p;
    rp.sums = c->d_sums.p;
    rp.semi_major = c->o.semi_major;
    rp.inv_flattening = c->o.inv_flattening;
    rp.critical = c->critical;
    rp.normals = normals;
    rp.assemble = assemble;
}

static void fill_clusters(gadj_ctx* c, ClusterParams& cp, int normals)
{
    cp.clusters = c->d_clusters.p;
    cp.nclusters = (uint32_t)c->clusters.size();
    cp.row_l = c->d_row_l.p;
    cp.row_a = c->d_row_a.p;
    cp.row_t = c->d_row_t.p;
    cp.cvinv = c->d_cvinv.p;
    cp.cstn = c->d_cstn.p;
    cp.inc_ptr = c->d_inc_ptr.p;
    cp.inc = c->d_inc.p;
    cp.pair_word = c->d_pair_word.p;
    cp.ndiag = c->d_ndiag.p;
    cp.noff = c->d_noff.p;
    cp.w = c->d_w.p;
    cp.sums = c->d_sums.p;
    cp.normals = normals;
}

static void fill_scatter(gadj_ctx* c, ScatterParams& sp)
{
    sp.ndiag = c->d_ndiag.p;
    sp.noff = c->d_noff.p;
    sp.bq = c->d_bq.p;
    sp.edge_bsl = c->d_edge_bsl.p;
    sp.diag_dest = c->d_diag_dest.p;
    sp.diag_ld = c->d_diag_ld.p;
    sp.off_dest = c->d_off_dest.p;
    sp.off_ld = c->d_off_ld.p;
    sp.edge_hi = c->d_edge_hi.p;
    sp.edge_lo = c->d_edge_lo.p;
    sp.panels = c->d_panels.p;
    sp.dscale = c->d_dscale.p;
    sp.nstn = c->nstn;
    sp.nedge = c->nedge;
    sp.scale = c->o.scale_normals_to_unity ? 1 : 0;
}

static int extract_vcv(gadj_ctx* c);

// ---- the stages of one iteration --------------------------------------------------------------------
// gadj_iterate runs them back to back.  In a multi-GPU run every rank executes the same sequence on its own
// stream; the launch lists carry the device-side barriers and all-reduces that join the ranks (plan.cpp).

static int stage_begin(gadj_ctx* c, int flags)
{
    if (!c->prepared)
        return c->fail("gadj_prepare has not been run");
    void* st = dev::stream();
    // non-GNSS rows: the partials move with the estimates, so the normals are rebuilt on every iteration and the
    // geographic coordinates are refreshed first (UpdateAdjustment, ADJ:543-545, 583-589)
    const bool normals = (flags & GADJ_ITER_NORMALS) || !c->factor_valid || c->non_gps;
    c->stage_normals = normals;
    dev::event_record(c->ev[0]);
    if (c->non_gps && c->iteration > 0)
        launch_cart_to_geo(c->d_est.p, c->d_llh.p, c->nstn, c->o.semi_major, c->o.inv_flattening, st);
    // ---- assembly (FillDesignNormalMeasurementsMatrices, ADJ:3888) --------------
    c->prof_begin(PK_ASSEMBLE);
    launch_init_normals(normals ? c->d_cblock.p : nullptr, c->d_ndiag.p, c->d_noff.p, c->d_w.p, c->nstn,
                        normals ? c->nedge : 0, st);
    c->prof_end();
    AssembleParams ap;
    fill_assemble(c, ap, normals ? 1 : 0);
    c->prof_begin(PK_ASSEMBLE);
    launch_assemble_g(ap, st);
    c->prof_end();
    c->prof_begin(PK_ASSEMBLE);
    launch_station_sum(ap, st);
    c->prof_end();
    if (c->nrows) {
        RowsParams rp;
        fill_rows(c, rp, normals ? 1 : 0, 1);
        c->prof_begin(PK_ASSEMBLE);
        launch_rows(rp, st);
        c->prof_end();
        if (!c->clusters.empty()) {
            ClusterParams cp;
            fill_clusters(c, cp, normals ? 1 : 0);
            c->prof_begin(PK_ASSEMBLE);
            launch_clusters(cp, st);
            c->prof_end();
        }
    }
    dev::event_record(c->ev[1]);
    // ---- equilibrate + scatter into the front panels ------------------------------
    if (normals) {
        c->inverse_valid = false;
        c->factor_valid = false;
        c->vcv_extracted = false;
        c->h_dscale_iteration = ~0u;   // new equilibration factors: the host copy is stale
        ScatterParams sp;
        fill_scatter(c, sp);
        c->prof_begin(PK_OTHER);
        launch_compute_scale(sp, st);
        c->prof_end();
        c->prof_begin(L_ZERO);
        c->launch_count--;
        dev::zero(c->d_panels.p, c->d_panels.bytes());
        dev::zero(c->d_info.p, c->d_info.bytes());
        c->prof_end();
        c->prof_begin(PK_OTHER);
        launch_scatter_normals(sp, st);
        c->prof_end();
        c->normals_valid = true;
    }
    return 0;
}

static int stage_solve_begin(gadj_ctx* c)
{
    dev::event_record(c->ev[2]);
    c->prof_begin(PK_OTHER);
    launch_permute_rhs(c->d_w.p, c->d_dscale.p, c->d_pos.p, c->mg_world > 1 ? c->d_pos_owned.p : nullptr, c->d_x.p, c->nstn,
                       dev::stream());
    c->prof_end();
    return 0;
}

// multi-GPU: every rank holds the solution of the replicated top fronts and of its own subtrees; the entries of the other
// ranks' subtrees are zeroed and the vector summed over the ranks (every rank updates all estimates: the next assembly
// is replicated)
static int stage_solve_end(gadj_ctx* c)
{
    if (c->mg_world <= 1)
        return 0;
    void* st = dev::stream();
    launch_mask_positions(c->d_x.p, c->d_pos_owned.p, c->nstn, st);
    mg_barrier(c);
    launch_allreduce(c->d_reduce_misc.p + 0, 1, c->pt(), c->d_x.p, MC_X, st);
    mg_barrier(c);
    // a non-positive pivot met by any rank (the owner of the pivot tile) is everybody's
    launch_share_info(c->pt(), c->d_info.p, st);
    mg_barrier(c);
    return 0;
}

static int stage_apply(gadj_ctx* c)
{
    c->prof_begin(PK_OTHER);
    c->launch_count++;  // two kernels: update + max reduction
    launch_apply_corrections(c->d_x.p, c->d_dscale.p, c->d_pos.p, c->d_corr.p, c->d_est.p, c->nstn, c->d_apply.p, dev::stream());
    c->prof_end();
    dev::event_record(c->ev[3]);
    return 0;
}

static int stage_end(gadj_ctx* c, int flags, gadj_iter_result* res)
{
    dev::event_record(c->ev[4]);
    int32_t info[4] = {0, 0, 0, 0};
    double tail[8];
    dev::d2h(info, c->d_info.p, sizeof(info));
    dev::d2h(tail, c->d_corr.p + 3 * (size_t)c->nstn, sizeof(tail));
    std::string e = dev::sync();
    if (!e.empty())
        return c->fail(e);
    if (info[1] != 0)
        return c->fail("multi-GPU barrier timed out: another rank has stopped");
    if (c->stage_normals) {
        if (info[0] != 0) {
            c->factor_valid = false;
            if (getenv("GADJ_DEBUG") && info[0] - 1 < (int)c->sym.fronts.size()) {
                const Front& f = c->sym.fronts[info[0] - 1];
                fprintf(stderr, "gadj: non-positive pivot in front %d (level %d, k=%u r=%u parent=%d)\n", info[0] - 1, f.level,
                        f.k, f.r, f.parent);
            }
            return c->fail("Matrix inversion failed, the matrix is singular.");
        }
        c->factor_valid = true;
    }
    if (flags & GADJ_ITER_INVERSE) {
        c->inverse_valid = true;
        c->vcv_extracted = false;
        c->factor_valid = false;  // the panels now hold the inverse, not the factor
    }
    c->iteration++;
    if (res) {
        std::memset(res, 0, sizeof(*res));
        res->max_corr = tail[0];
        uint64_t row = (uint64_t)tail[1];
        res->max_corr_station = (uint32_t)(row / 3);
        res->max_corr_axis = (uint32_t)(row % 3);
        res->iteration = c->iteration;
        res->converged = std::fabs(tail[0]) <= c->o.iteration_threshold;
        res->ms_assemble = dev::event_elapsed_ms(c->ev[0], c->ev[1]);
        res->ms_factor = dev::event_elapsed_ms(c->ev[1], c->ev[2]);
        res->ms_solve = dev::event_elapsed_ms(c->ev[2], c->ev[3]);
        res->ms_inverse = dev::event_elapsed_ms(c->ev[3], c->ev[4]);
        res->max_corr_xyz[0] = tail[2];
        res->max_corr_xyz[1] = tail[3];
        res->max_corr_xyz[2] = tail[4];
        if (std::isnan(tail[0]) || std::isinf(tail[0]))
            return c->fail("Solve(): Invalid variance matrix");
    }
    return 0;
}

int gadj_iterate(gadj_ctx* c, int flags, gadj_iter_result* res)
{
    dev::use(c->device);
    if (!c->prepared)
        return c->fail("gadj_prepare has not been run");
    if (c->mg_world > 1 && !c->connected)
        return c->fail("this context is one rank of a multi-GPU adjustment: exchange the peer handles first (gadj_mg_connect)");
    if (stage_begin(c, flags))
        return 1;
    if (c->stage_normals)
        run_launches(c, c->plan.factor);
    stage_solve_begin(c);
    run_launches(c, c->plan.fwd);
    run_launches(c, c->plan.bwd);
    stage_solve_end(c);
    stage_apply(c);
    if (flags & GADJ_ITER_INVERSE)
        run_launches(c, c->plan.selinv);
    return stage_end(c, flags, res);
}

int gadj_sync(gadj_ctx* c)
{
    dev::use(c->device);
    std::string e = dev::sync();
    return e.empty() ? 0 : c->fail(e);
}

int gadj_mg_init(gadj_ctx* c, int32_t rank, int32_t world)
{
    dev::use(c->device);
    if (world < 1 || rank < 0 || rank >= world)
        return c->fail("bad rank / world size");
    c->mg_rank = rank;
    c->mg_world = world;
    c->prepared = false;
    return 0;
}

int gadj_mg_buffer(gadj_ctx* c, int which, void** ptr, uint64_t* count)
{
    dev::use(c->device);
    if (!c->prepared)
        return c->fail("gadj_prepare has not been run");
    switch (which) {
    case 0:
        *ptr = c->d_x.p;
        *count = c->d_x.n;
        break;
    case 1:
        *ptr = c->d_panels.p;
        *count = c->d_panels.n;
        break;
    case 2:
        *ptr = c->d_vcvd.p;
        *count = c->d_vcvd.n;
        break;
    case 3:
        *ptr = c->d_vcvo.p;
        *count = c->d_vcvo.n;
        break;
    case 4:
        *ptr = c->d_info.p;
        *count = c->d_info.n;
        break;
    case 5:   // the device copy of the measurement records, in bytes
        *ptr = c->d_msr.p;
        *count = c->d_msr.n * sizeof(dna_msr_t);
        break;
    case 6:   // W = L11^-1 and its transpose of every owned front (diagnostics)
        *ptr = c->d_wbuf.p;
        *count = c->d_wbuf.n;
        break;
    case 7:   // selected-inverse workspace pool (diagnostics)
        *ptr = c->d_pool.p;
        *count = c->d_pool.n;
        break;
    default:
        return c->fail("unknown buffer");
    }
    return 0;
}

// ---- multi-GPU: one rank per GPU, peers' buffers mapped over NVLink --------------------------------------------
namespace {
struct ExportedBuffer {
    int which;
    void* p;
    size_t bytes;
};
// the peer-visible buffers of a prepared context, in gadj_peer_info order (index = McBuf, 0 = the barrier counter)
void exported_buffers(gadj_ctx* c, ExportedBuffer out[GADJ_PEER_BUFFERS])
{
    for (int i = 0; i < GADJ_PEER_BUFFERS; ++i)
        out[i] = ExportedBuffer{i, nullptr, 0};
    out[0] = {0, c->d_counter.p, c->d_counter.bytes()};
    out[MC_PANELS] = {MC_PANELS, c->d_panels.p, c->d_panels.bytes()};
    out[MC_WBUF] = {MC_WBUF, c->d_wbuf.p, c->d_wbuf.bytes()};
    out[MC_POOL] = {MC_POOL, c->d_pool.p, c->d_pool.bytes()};
    out[MC_X] = {MC_X, c->d_x.p, c->d_x.bytes()};
    out[MC_VCVD] = {MC_VCVD, c->d_vcvd.p, c->d_vcvd.bytes()};
    out[MC_VCVO] = {MC_VCVO, c->d_vcvo.p, c->d_vcvo.bytes()};
    out[MC_MSR] = {MC_MSR, c->d_msr.p, c->d_msr.bytes()};
    out[MC_BUFS] = {MC_BUFS, c->d_info.p, c->d_info.bytes()};
}
}  // namespace

int gadj_mg_export(gadj_ctx* c, gadj_peer_info* out)
{
    dev::use(c->device);
    if (!c->prepared)
        return c->fail("gadj_prepare has not been run");
    if (c->mg_world <= 1)
        return c->fail("not a multi-GPU context (gadj_mg_init)");
    static_assert(GADJ_PEER_BUFFERS == MC_BUFS + 1, "peer buffer table");
    static_assert(GADJ_IPC_HANDLE_BYTES == dev::IPC_HANDLE_BYTES, "IPC handle size");
    std::memset(out, 0, sizeof(*out));
    out->rank = c->mg_rank;
    out->device = dev::ordinal();
    out->pid = dev::process_id();
    ExportedBuffer eb[GADJ_PEER_BUFFERS];
    exported_buffers(c, eb);
    for (int i = 0; i < GADJ_PEER_BUFFERS; ++i) {
        out->ptr[i] = (uint64_t)(uintptr_t)eb[i].p;
        out->bytes[i] = eb[i].bytes;
        if (eb[i].p && !dev::ipc_export(eb[i].p, eb[i].bytes, out->handle[i]))
            return c->fail("cannot export a device buffer to the other ranks (IPC handle)");
    }
    // the replicated part of the layout must be the same on every rank
    out->top_panel_doubles = c->sym.top_panel_doubles;
    out->nstations = c->nstn;
    return 0;
}

int gadj_mg_connect(gadj_ctx* c, const gadj_peer_info* all)
{
    dev::use(c->device);
    if (!c->prepared)
        return c->fail("gadj_prepare has not been run");
    if (c->mg_world <= 1)
        return c->fail("not a multi-GPU context (gadj_mg_init)");
    if (c->mg_world > MAX_RANKS)
        return c->fail("at most " + std::to_string(MAX_RANKS) + " ranks (the GPUs of one NVLink node)");
    ExportedBuffer mine[GADJ_PEER_BUFFERS];
    exported_buffers(c, mine);
    PeerTable& t = c->peers;
    std::memset(&t, 0, sizeof(t));
    t.nranks = c->mg_world;
    t.rank = c->mg_rank;
    for (int q = 0; q < c->mg_world; ++q) {
        const gadj_peer_info& pi = all[q];
        if (pi.rank != q)
            return c->fail("peer table out of order");
        if (pi.top_panel_doubles != c->sym.top_panel_doubles || pi.nstations != c->nstn)
            return c->fail("the ranks prepared different networks / orderings");
        for (int i = 0; i < GADJ_PEER_BUFFERS; ++i) {
            void* mapped = nullptr;
            if (q == c->mg_rank)
                mapped = mine[i].p;
            else if (pi.ptr[i]) {
                std::string e;
                mapped = dev::peer_map(pi.device, pi.pid, (void*)(uintptr_t)pi.ptr[i], pi.handle[i], pi.bytes[i], e);
                if (!mapped)
                    return c->fail("cannot map a buffer of rank " + std::to_string(q) + ": " + e);
                c->peer_maps.push_back({mapped, pi.pid});
            }
            if (i == 0)
                t.counter[q] = (unsigned long long*)mapped;
            else if (i == MC_BUFS)
                t.info[q] = (int32_t*)mapped;
            else
                t.delta[i][q] = mapped && mine[i].p ? (int64_t)((char*)mapped - (char*)mine[i].p) : 0;
        }
    }
    if (!c->d_peers.resize(1))
        return c->fail("out of device memory");
    dev::h2d(c->d_peers.p, &t, sizeof(t));
    std::string e = dev::sync();
    if (!e.empty())
        return c->fail(e);
    c->connected = true;
    return 0;
}

int gadj_form_inverse(gadj_ctx* c)
{
    dev::use(c->device);
    if (!c->prepared)
        return c->fail("gadj_prepare has not been run");
    if (c->inverse_valid)
        return 0;
    if (!c->factor_valid)
        return c->fail("no valid factorisation to invert (run gadj_iterate first)");
    if (c->mg_world > 1 && !c->connected)
        return c->fail("this context is one rank of a multi-GPU adjustment: exchange the peer handles first (gadj_mg_connect)");
    dev::event_record(c->ev[3]);
    run_launches(c, c->plan.selinv);
    dev::event_record(c->ev[4]);
    std::string e = dev::sync();
    if (!e.empty())
        return c->fail(e);
    c->inverse_valid = true;
    c->vcv_extracted = false;
    c->factor_valid = false;
    return 0;
}

int gadj_adjust(gadj_ctx* c, gadj_iter_result* last)
{
    dev::use(c->device);
    if (!c->prepared)
        return c->fail("gadj_prepare has not been run");
    gadj_iter_result r{};
    // AdjustSimultaneous (ADJ:2413-2511): the normals of a GNSS-only network never change, so
    // iterations >= 2 re-use the factorisation exactly as the reference re-uses its inverse.
    float ms_inv = 0;
    for (uint32_t i = 0; i < c->o.max_iterations; ++i) {
        int flags = (c->iteration < 1) ? GADJ_ITER_NORMALS : 0;
        if (gadj_iterate(c, flags, &r))
            return 1;
        if (std::fabs(r.max_corr) <= c->o.iteration_threshold)
            break;
    }
    // rigorous variances (the reference carries them in v_normals_ after Solve)
    if (!c->inverse_valid) {
        if (gadj_form_inverse(c))
            return 1;
        ms_inv = dev::event_elapsed_ms(c->ev[3], c->ev[4]);
    }
    r.ms_inverse = ms_inv;
    if (last)
        *last = r;
    return 0;
}

static int extract_vcv(gadj_ctx* c)
{
    if (!c->inverse_valid)
        return c->fail("the rigorous inverse has not been formed (run gadj_adjust or iterate with GADJ_ITER_INVERSE)");
    if (c->vcv_extracted)
        return 0;
    c->vcv_extracted = true;
    void* st = dev::stream();
    launch_extract_station_vcv(c->d_panels.p, c->d_diag_dest.p, c->d_diag_ld.p, c->d_dscale.p, c->d_vcvd.p, c->nstn, st);
    launch_extract_edge_vcv(c->d_panels.p, c->d_off_dest.p, c->d_off_ld.p, c->d_edge_hi.p, c->d_edge_lo.p, c->d_dscale.p,
                            c->d_vcvo.p, c->nedge, st);
    if (c->mg_world > 1) {
        // every rank extracted the blocks of the fronts it assembles (zeros elsewhere): the sum over the ranks is the whole
        mg_barrier(c);
        launch_allreduce(c->d_reduce_misc.p + 1, 1, c->pt(), c->d_vcvd.p, MC_VCVD, st);
        if (c->nedge)
            launch_allreduce(c->d_reduce_misc.p + 2, 1, c->pt(), c->d_vcvo.p, MC_VCVO, st);
        mg_barrier(c);
    }
    return 0;
}

int gadj_statistics(gadj_ctx* c, gadj_stats* stt, int write_back)
{
    dev::use(c->device);
    if (!c->prepared)
        return c->fail("gadj_prepare has not been run");
    if (extract_vcv(c))
        return 1;
    void* st = dev::stream();
    // UpdateAdjustment(false): geographic coordinates + re-linearised l (ADJ:6807, ADJ:8734)
    launch_cart_to_geo(c->d_est.p, c->d_llh.p, c->nstn, c->o.semi_major, c->o.inv_flattening, st);
    dev::zero(c->d_sums.p, c->d_sums.bytes());
    StatsParams sp;
    sp.msr = c->d_msr.p;
    sp.first = c->d_first.p;
    sp.edge = c->d_edge.p;
    sp.est = c->d_est.p;
    sp.vcv_diag = c->d_vcvd.p;
    sp.vcv_off = c->d_vcvo.p;
    sp.sums = c->d_sums.p;
    sp.nbaselines = c->nbsl;
    sp.critical = c->critical;
    launch_stats_g(sp, st);
    if (c->nrows) {
        // re-linearise the rows at the final estimates, then per-row statistics and the cluster chi-squares
        RowsParams rp;
        fill_rows(c, rp, 0, 0);
        launch_rows(rp, st);
        launch_rows_stats(rp, st);
        if (!c->clusters.empty()) {
            ClusterParams cp;
            fill_clusters(c, cp, 0);
            launch_cluster_chi(cp, st);
        }
    }
    double sums[8];
    dev::d2h(sums, c->d_sums.p, sizeof(sums));
    std::vector<double> llh;
    if (write_back) {
        dev::d2h(c->msr, c->d_msr.p, c->nmsr * sizeof(dna_msr_t));
        llh.resize(3 * (size_t)c->nstn);
        dev::d2h(llh.data(), c->d_llh.p, llh.size() * sizeof(double));
    }
    std::string e = dev::sync();
    if (!e.empty())
        return c->fail(e);
    if (write_back)
        for (uint32_t s = 0; s < c->nstn; ++s) {
            c->stn[s].currentLatitude = llh[3 * s];
            c->stn[s].currentLongitude = llh[3 * s + 1];
            c->stn[s].currentHeight = llh[3 * s + 2];
        }
    std::memset(stt, 0, sizeof(*stt));
    stt->chi_squared = sums[0];
    stt->measurement_params = (uint32_t)(3 * c->nbsl + c->nrows);
    stt->unknown_params = 3 * (c->nstn - c->unused_stations) - c->constrained_components;
    stt->dof = (int64_t)stt->measurement_params - (int64_t)stt->unknown_params;   // ADJ:6856
    stt->sigma_zero = stt->dof != 0 ? stt->chi_squared / (double)stt->dof : 0.0;
    stt->outliers = (uint32_t)sums[3];
    stt->global_pelzer = sums[2] > 0 ? std::sqrt(sums[1] / sums[2]) : 999.99;       // ADJ:8350-8354
    stt->critical_value = c->critical;
    return 0;
}

// ---- ignored measurements, a posteriori (UpdateIgnoredMeasurements_* ADJ:8750-9980; reporting only, host side) -------
// For every measurement flagged as ignored whose stations were adjusted: the value computed from the adjusted
// coordinates, mapped back to the domain the measurement was observed in (deflection / geoid / arc reductions undone),
// and its difference from the measured value.  Writes preAdjMeas, measAdj, measCorr, preAdjCorr of the ignored records
// (direction sets: scale1 = derived angle, scale2 = its variance, on the direction records).  Call after gadj_statistics
// with write_back (the stations' geographic coordinates are then the adjusted ones).
static int compute_measurements_host(gadj_ctx* c, bool want_ignored)
{
    dev::use(c->device);
    if (!c->prepared)
        return c->fail("gadj_prepare has not been run");
    const uint32_t ns = c->nstn;
    std::vector<double> est(3 * (size_t)ns), llh(3 * (size_t)ns);
    std::vector<float> geoid(ns);
    dev::d2h(est.data(), c->d_est.p, est.size() * sizeof(double));
    std::string e = dev::sync();
    if (!e.empty())
        return c->fail(e);
    for (uint32_t s = 0; s < ns; ++s) {
        cart_to_geo(c->ell, est[3 * s], est[3 * s + 1], est[3 * s + 2], &llh[3 * (size_t)s]);
        geoid[s] = c->stn[s].geoidSep;
    }
    auto bad = [&](uint32_t s) { return s >= ns; };
    // one scalar row (also the derived angle of a direction set): geometric value -> observation domain (ADJ:8194-8271)
    auto scalar = [&](dna_msr_t& m, char type, const uint32_t st[3], int nst) {
        RowDesc d;
        std::memset(&d, 0, sizeof(d));
        d.type = (uint8_t)type;
        d.nst = (uint8_t)nst;
        for (int k = 0; k < 3; ++k)
            d.st[k] = st[k];
        if (want_ignored || !c->reduced || type == 'D')
            m.preAdjMeas = m.term1;           // else: reduced by an earlier run, the measured value is kept in preAdjMeas
        dna_msr_t w = m;                      // the reductions act on a copy: the record keeps its value
        w.term1 = m.preAdjMeas;
        first_run_reduce(type, w, d.st, c->stn, est.data());
        RowOut o;
        double reduced = 0.0;
        if (!design_row(d, w.term1, w.term3, w.term4, w.preAdjMeas, est.data(), llh.data(), geoid.data(), c->ell, o, &reduced))
            return;
        if (type == 'E' || type == 'M') {
            w.term1 = reduced;
            w.preAdjCorr = reduced - w.preAdjMeas;
        }
        double adj = w.term1 - o.l;           // the value computed from the adjusted coordinates, as reduced
        switch (type) {
        case 'D':
            if (adj > kTwoPi)
                adj -= kTwoPi;
            adj += w.preAdjCorr;
            break;
        case 'E': {
            const double r = chord_radius(c->ell, &est[3 * (size_t)st[0]], &est[3 * (size_t)st[1]], &llh[3 * (size_t)st[0]], &llh[3 * (size_t)st[1]]);
            adj = std::asin(adj / 2.0 / r) * 2.0 * r;
            break;
        }
        case 'M':
            adj = chord_to_msl_arc(c->ell, adj, llh[3 * (size_t)st[0]], llh[3 * (size_t)st[1]], (double)geoid[st[0]], (double)geoid[st[1]]);
            break;
        case 'H': case 'L': case 'V':
            adj -= w.preAdjCorr;
            break;
        case 'A': case 'I': case 'J': case 'K': case 'Z':
            adj += w.preAdjCorr;
            break;
        default:
            break;
        }
        m.preAdjCorr = w.preAdjCorr;
        m.measAdj = adj;
        m.measCorr = adj - m.preAdjMeas;
        if (std::strchr("ABDKVZ", type)) {   // angular differences stay within half a turn
            if (m.measCorr > kPi)
                m.measCorr -= kTwoPi;
            if (m.measCorr < -kPi)
                m.measCorr += kTwoPi;
        }
    };
    uint64_t i = 0;
    while (i < c->nmsr) {
        dna_msr_t& m = c->msr[i];
        uint64_t step = 1;
        switch (m.measType) {
        case 'G':
            step = 3;
            break;
        case 'X': case 'Y': {
            uint64_t j = i;
            for (uint32_t k = 0; k < std::max<uint32_t>(1u, m.vectorCount1) && j < c->nmsr; ++k)
                j += 3 + 3ull * c->msr[j].vectorCount2;
            step = j - i;
            break;
        }
        case 'D':
            step = std::max<uint32_t>(1u, m.vectorCount1);
            break;
        default:
            break;
        }
        if (i + step > c->nmsr)
            break;
        if ((m.ignore != 0) == want_ignored) {
            if (m.measType == 'G' || m.measType == 'X' || m.measType == 'Y') {
                const bool geographic = m.measType == 'Y' && (std::strncmp(m.coordType, "LLH", 3) == 0 || std::strncmp(m.coordType, "LLh", 3) == 0);
                for (uint64_t j = i; j + 2 < i + step;) {
                    dna_msr_t* r = &c->msr[j];
                    const uint32_t s1 = r->station1, s2 = m.measType == 'Y' ? r->station1 : r->station2;
                    if (!bad(s1) && !bad(s2)) {
                        double v[3];
                        for (int q = 0; q < 3; ++q)
                            v[q] = m.measType == 'Y' ? est[3 * (size_t)s1 + q] : est[3 * (size_t)s2 + q] - est[3 * (size_t)s1 + q];
                        if (geographic) {       // UpdateIgnoredMeasurements_Y (ADJ:9720-9830): latitude, longitude, height as supplied
                            v[0] = llh[3 * (size_t)s1], v[1] = llh[3 * (size_t)s1 + 1], v[2] = llh[3 * (size_t)s1 + 2];
                            if (std::strncmp(m.coordType, "LLH", 3) == 0 && std::fabs((double)geoid[s1]) > 1.0e-4) {
                                r[2].preAdjCorr = geoid[s1];
                                v[2] -= geoid[s1];
                            }
                        }
                        for (int q = 0; q < 3; ++q) {
                            if (want_ignored || !c->reduced)
                                r[q].preAdjMeas = r[q].term1;
                            r[q].measAdj = v[q];
                            r[q].measCorr = v[q] - r[q].preAdjMeas;
                        }
                    }
                    j += 3 + 3ull * r->vectorCount2;
                }
            } else if (m.measType == 'D') {
                double prev_dir = m.term1, prev_var = m.term2;
                uint64_t prev = i;
                for (uint64_t j = i + 1; j < i + step; ++j) {
                    dna_msr_t& dir = c->msr[j];
                    if (!want_ignored && dir.ignore)
                        continue;             // an ignored direction inside a set that is used (ADJ:5120-5129)
                    const uint32_t st[3] = {c->msr[prev].station1, c->msr[prev].station2, dir.station2};
                    if (!bad(st[0]) && !bad(st[1]) && !bad(st[2]) && st[0] != st[1] && st[0] != st[2] && st[1] != st[2]) {
                        dna_msr_t angle = c->msr[prev];
                        angle.term1 = dir.term1 - prev_dir;
                        if (angle.term1 < 0)
                            angle.term1 += kTwoPi;
                        if (angle.term1 > kTwoPi)
                            angle.term1 -= kTwoPi;
                        scalar(angle, 'D', st, 3);
                        dir.scale1 = angle.preAdjMeas;
                        dir.scale2 = prev_var + dir.term2;
                        dir.measCorr = angle.measCorr;
                        dir.measAdj = angle.measAdj;
                        dir.preAdjCorr = angle.preAdjCorr;
                    }
                    prev_dir = dir.term1;
                    prev_var = dir.term2;
                    prev = j;
                }
            } else if (is_scalar_type(m.measType)) {
                const int nst = stations_of_type(m.measType);
                const uint32_t st[3] = {m.station1, nst > 1 ? m.station2 : m.station1, nst > 2 ? m.station3 : m.station1};
                if (!bad(st[0]) && !bad(st[1]) && !bad(st[2]) && (nst < 2 || st[0] != st[1]) && (nst < 3 || (st[0] != st[2] && st[1] != st[2])))
                    scalar(m, m.measType, st, nst);
            }
        }
        i += step;
    }
    return 0;
}

int gadj_update_ignored_measurements(gadj_ctx* c) { return compute_measurements_host(c, true); }

// "Computed Measurements (a-priori)" of --output-iter-cmp-msr (PrintCompMeasurements PRN:1938-2023): the same evaluation
// for the measurements that take part, at the current estimates.  Meant for the start of the first iteration — later
// iterations read the re-linearised records of gadj_statistics.
int gadj_compute_measurements(gadj_ctx* c) { return compute_measurements_host(c, false); }

int gadj_get_estimates(gadj_ctx* c, double* xyz)
{
    dev::use(c->device);
    if (!c->prepared)
        return c->fail("gadj_prepare has not been run");
    dev::d2h(xyz, c->d_est.p, c->d_est.bytes());
    std::string e = dev::sync();
    return e.empty() ? 0 : c->fail(e);
}

int gadj_get_corrections(gadj_ctx* c, double* dxyz)
{
    dev::use(c->device);
    if (!c->prepared)
        return c->fail("gadj_prepare has not been run");
    dev::d2h(dxyz, c->d_corr.p, 3 * (size_t)c->nstn * sizeof(double));
    std::string e = dev::sync();
    return e.empty() ? 0 : c->fail(e);
}

int gadj_get_station_vcvs(gadj_ctx* c, double* q)
{
    dev::use(c->device);
    if (!c->prepared)
        return c->fail("gadj_prepare has not been run");
    if (extract_vcv(c))
        return 1;
    dev::d2h(q, c->d_vcvd.p, c->d_vcvd.bytes());
    std::string e = dev::sync();
    return e.empty() ? 0 : c->fail(e);
}

int gadj_get_station_vcv(gadj_ctx* c, uint32_t stn, double q[9])
{
    dev::use(c->device);
    if (!c->prepared)
        return c->fail("gadj_prepare has not been run");
    if (stn >= c->nstn)
        return c->fail("station index out of range");
    if (extract_vcv(c))
        return 1;
    dev::d2h(q, c->d_vcvd.p + 9 * (size_t)stn, 9 * sizeof(double));
    std::string e = dev::sync();
    return e.empty() ? 0 : c->fail(e);
}

static int find_edge(gadj_ctx* c, uint32_t si, uint32_t sj, uint64_t* ei, bool* transposed)
{
    // edges are sorted by (min station, max station); edge_hi/edge_lo carry the elimination orientation
    uint32_t lo = std::min(si, sj), hi = std::max(si, sj);
    uint64_t b = 0, e = c->nedge;
    while (b < e) {
        uint64_t mid = (b + e) / 2;
        uint32_t mlo = std::min(c->edge_hi[mid], c->edge_lo[mid]), mhi = std::max(c->edge_hi[mid], c->edge_lo[mid]);
        if (mlo < lo || (mlo == lo && mhi < hi))
            b = mid + 1;
        else
            e = mid;
    }
    if (b >= c->nedge)
        return 1;
    uint32_t mlo = std::min(c->edge_hi[b], c->edge_lo[b]), mhi = std::max(c->edge_hi[b], c->edge_lo[b]);
    if (mlo != lo || mhi != hi)
        return 1;
    *ei = b;
    *transposed = (c->edge_hi[b] != si);  // stored block is N[hi, lo]
    return 0;
}

static int get_block(gadj_ctx* c, const double* diag, const double* off, uint32_t si, uint32_t sj, double q[9])
{
    if (si >= c->nstn || sj >= c->nstn)
        return c->fail("station index out of range");
    double t[9];
    bool tr = false;
    if (si == sj)
        dev::d2h(t, diag + 9 * (size_t)si, sizeof(t));
    else {
        uint64_t ei;
        if (find_edge(c, si, sj, &ei, &tr))
            return c->fail("no measurement joins the two stations: block is outside the stored pattern");
        dev::d2h(t, off + 9 * ei, sizeof(t));
    }
    std::string e = dev::sync();
    if (!e.empty())
        return c->fail(e);
    for (int r = 0; r < 3; ++r)
        for (int k = 0; k < 3; ++k)
            q[r * 3 + k] = tr ? t[k * 3 + r] : t[r * 3 + k];
    return 0;
}

int gadj_get_vcv_block(gadj_ctx* c, uint32_t si, uint32_t sj, double q[9])
{
    dev::use(c->device);
    if (!c->prepared)
        return c->fail("gadj_prepare has not been run");
    if (extract_vcv(c))
        return 1;
    return get_block(c, c->d_vcvd.p, c->d_vcvo.p, si, sj, q);
}

// bulk form of gadj_get_vcv_block: one device -> host copy of the stored variance blocks, then the pairs on the host
int gadj_get_pair_vcvs(gadj_ctx* c, uint64_t npairs, const uint32_t* si, const uint32_t* sj, double* q)
{
    dev::use(c->device);
    if (!c->prepared)
        return c->fail("gadj_prepare has not been run");
    if (extract_vcv(c))
        return 1;
    std::vector<double> diag(9 * (size_t)c->nstn), off(9 * (size_t)c->nedge);
    dev::d2h(diag.data(), c->d_vcvd.p, diag.size() * sizeof(double));
    if (!off.empty())
        dev::d2h(off.data(), c->d_vcvo.p, off.size() * sizeof(double));
    std::string e = dev::sync();
    if (!e.empty())
        return c->fail(e);
    for (uint64_t p = 0; p < npairs; ++p) {
        if (si[p] >= c->nstn || sj[p] >= c->nstn)
            return c->fail("station index out of range");
        const double* t;
        bool tr = false;
        if (si[p] == sj[p])
            t = &diag[9 * (size_t)si[p]];
        else {
            uint64_t ei;
            if (find_edge(c, si[p], sj[p], &ei, &tr))
                return c->fail("no measurement joins the two stations: block is outside the stored pattern");
            t = &off[9 * ei];
        }
        for (int r = 0; r < 3; ++r)
            for (int k = 0; k < 3; ++k)
                q[9 * p + r * 3 + k] = tr ? t[k * 3 + r] : t[r * 3 + k];
    }
    return 0;
}

int gadj_get_block_vcv(gadj_ctx* c, uint32_t block, uint32_t* nstations, uint32_t* stations, uint32_t cap, double* packed_lower)
{
    dev::use(c->device);
    if (!c->prepared)
        return c->fail("gadj_prepare has not been run");
    if (!c->inverse_valid)
        return c->fail("the rigorous variances have not been formed");
    const Symbolic& S = c->sym;
    if (block >= S.fronts.size())
        return c->fail("block index out of range");
    const Front& f = S.fronts[block];
    if (f.owner != S.rank && !f.top)
        return c->fail("the block is held by another rank");
    const uint32_t nst = f.own_count + f.bnd_count;
    *nstations = nst;
    if (!stations || !packed_lower)
        return 0;
    if (cap < nst)
        return c->fail("station list capacity too small");
    for (uint32_t i = 0; i < f.own_count; ++i)
        stations[i] = S.stn_of_pos[f.own_begin + i];
    for (uint32_t i = 0; i < f.bnd_count; ++i)
        stations[f.own_count + i] = S.stn_of_pos[S.bnd[f.bnd_begin + i]];
    const size_t k = f.k, r = f.r, m = f.m, n = m;
    // the front's panel [Z11; Z21] and the junction block Z22, gathered from the ancestors' panels like the selected inverse does
    std::vector<double> P(m * (size_t)f.ldk), G;
    dev::d2h(P.data(), c->d_panels.p + f.panel_off, P.size() * sizeof(double));
    // the equilibration factors of the inverse in the panels: fetched once per inverse, not once per block
    if (c->h_dscale_iteration != c->iteration + 1 || c->h_dscale.size() != 3 * (size_t)c->nstn) {
        c->h_dscale.resize(3 * (size_t)c->nstn);
        dev::d2h(c->h_dscale.data(), c->d_dscale.p, c->h_dscale.size() * sizeof(double));
        c->h_dscale_iteration = c->iteration + 1;
    }
    const std::vector<double>& d = c->h_dscale;
    const size_t ldg = r + (r & 1);
    DevArray<double> dG;
    DevArray<GatherOp> dops;
    DevArray<GatherTile> dtiles;
    if (r) {
        std::vector<GatherOp> ops;
        if (!dG.resize(r * ldg))
            return c->fail("out of device memory");
        for (uint32_t t = 0; t < f.tgt_count; ++t) {
            const Target& tg = S.targets[f.tgt_begin + t];
            const Front& an = S.fronts[tg.anc];
            GatherOp g{};
            g.Z = c->d_panels.p + an.panel_off;
            g.ld = an.ldk;
            g.rowmap = c->d_rowmap.p + tg.rowmap_off;
            g.G = dG.p;
            g.ldg = (int64_t)ldg;
            g.jb = (int32_t)tg.jb;
            g.je = (int32_t)tg.je;
            g.nb = (int32_t)f.bnd_count;
            g.col0 = (int32_t)tg.col0;
            ops.push_back(g);
        }
        std::vector<GatherTile> tiles;
        append_gather_tiles(ops, tiles);
        if (!dops.upload(ops) || !dtiles.upload(tiles))
            return c->fail("out of device memory");
        launch_gather(dops.p, (int)ops.size(), dtiles.p, (int)tiles.size(), dev::stream());
        G.resize(r * ldg);
        dev::d2h(G.data(), dG.p, G.size() * sizeof(double));
    }
    std::string e = dev::sync();
    if (!e.empty())
        return c->fail(e);
    auto scale_of = [&](size_t row) { return d[3 * (size_t)stations[row / 3] + row % 3]; };
    for (size_t j = 0; j < n; ++j)
        for (size_t i = j; i < n; ++i) {
            double z;
            if (j < k)
                z = P[i * f.ldk + j];                    // Z11 (lower triangle) and Z21
            else
                z = G[(i - k) * ldg + (j - k)];          // Z22
            packed_lower[j * n - j * (j - 1) / 2 + (i - j)] = z * scale_of(i) * scale_of(j);
        }
    return 0;
}

int gadj_get_normals_block(gadj_ctx* c, uint32_t si, uint32_t sj, double n[9])
{
    dev::use(c->device);
    if (!c->prepared || !c->normals_valid)
        return c->fail("normals have not been assembled");
    if (si != sj && si < c->nstn && sj < c->nstn) {
        uint64_t ei;
        bool tr = false;
        if (!find_edge(c, si, sj, &ei, &tr) && c->edge_bsl[ei] != ~0u) {
            // the pair is observed by a single GNSS baseline: its block -V^-1 lives in the baseline's slot
            double q6[6];
            dev::d2h(q6, c->d_bq.p + 9 * (size_t)c->edge_bsl[ei], sizeof(q6));
            std::string e = dev::sync();
            if (!e.empty())
                return c->fail(e);
            for (int k = 0; k < 9; ++k)
                n[k] = -q6[GADJ_SYM3(k)];
            return 0;
        }
    }
    return get_block(c, c->d_ndiag.p, c->d_noff.p, si, sj, n);
}

int gadj_get_rhs(gadj_ctx* c, double* w)
{
    dev::use(c->device);
    if (!c->prepared)
        return c->fail("gadj_prepare has not been run");
    dev::d2h(w, c->d_w.p, c->d_w.bytes());
    std::string e = dev::sync();
    return e.empty() ? 0 : c->fail(e);
}

int gadj_profile_enable(gadj_ctx* c, int on)
{
    dev::use(c->device);
    c->profiling = on != 0;
    return 0;
}

int gadj_profile_read(gadj_ctx* c, gadj_profile* out, int reset)
{
    dev::use(c->device);
    std::string e = dev::sync();
    if (!e.empty())
        return c->fail(e);
    FILE* dump = nullptr;
    if (const char* dp = getenv("GADJ_PROFILE_DUMP")) {
        std::string path = dp;
        if (c->mg_world > 1)
            path += ".rank" + std::to_string(c->mg_rank);
        dump = fopen(path.c_str(), "a");
    }
    for (size_t i = 0; i < c->prof_items.size(); ++i) {
        double ms = dev::event_elapsed_ms(c->prof_ev[2 * i], c->prof_ev[2 * i + 1]);
        const auto& it = c->prof_items[i];
        if (dump)
            fprintf(dump, "%zu,%d,%.6g,%d,%.6f,%d,%d\n", i, it.kind, it.flops, it.tiles, ms, it.tag, it.level);
        switch (it.kind) {
        case L_GEMM:
            c->prof.ms_gemm += ms;
            c->prof.flops_gemm += it.flops;
            c->prof.gemm_launches++;
            c->prof.gemm_tiles += (uint64_t)it.tiles;
            break;
        case L_DIAG:
            c->prof.ms_diag += ms;
            break;
        case L_TRI_FWD:
        case L_TRI_BWD:
            c->prof.ms_tri += ms;
            break;
        case L_GEMV_FWD:
        case L_GEMV_BWD:
            c->prof.ms_gemv += ms;
            break;
        case L_TRANSPOSE:
            c->prof.ms_transpose += ms;
            break;
        case L_GATHER:
            c->prof.ms_gather += ms;
            break;
        case L_ZERO:
            c->prof.ms_zero += ms;
            break;
        case PK_ASSEMBLE:
            c->prof.ms_assemble += ms;
            break;
        default:
            c->prof.ms_other += ms;
        }
    }
    if (dump)
        fclose(dump);
    c->prof_items.clear();
    c->prof.launches = c->launch_count;
    if (out)
        *out = c->prof;
    if (reset) {
        c->prof = gadj_profile{};
        c->launch_count = 0;
    }
    return 0;
}

int gadj_test_gemm_ex(gadj_ctx* c, const double* A, const double* B, double* C, double* Ct, int M, int N, int K, int flags,
                      int tile, int reps, float* ms)
{
    dev::use(c->device);
    if (M <= 0 || N <= 0 || K <= 0)
        return c->fail("gadj_test_gemm: M, N, K must be positive");
    if (tile != 64 && tile != 128)
        return c->fail("gadj_test_gemm: tile must be 64 or 128");
    const int allowed = GEMM_ACCUM | GEMM_NEG | GEMM_LOWER | GEMM_KLO_ROW | GEMM_KLO_MAX | GEMM_KHI_ROW | GEMM_DUAL;
    if ((flags & ~allowed) || ((flags & GEMM_DUAL) && (!Ct || (flags & (GEMM_ACCUM | GEMM_LOWER)))))
        return c->fail("gadj_test_gemm: unsupported flags");
    const int shape = tile == 64 ? TILE_SHAPE_64 : TILE_SHAPE_128;
    DevArray<double> dA, dB, dC, dCt;
    DevArray<GemmOp> dop;
    DevArray<GemmTile> dtl;
    // rows padded to an even pitch like every panel of the engine (odd K = 3 x an odd station count is the common case);
    // the padding holds NaNs: it lies beyond the tensor map's extent and must never reach the product
    const int ldc = N + (N & 1), ldk = K + (K & 1), ldct = M + (M & 1);
    if (!dA.resize((size_t)M * ldk) || !dB.resize((size_t)N * ldk) || !dC.resize((size_t)M * ldc) || !dop.resize(1) ||
        !dCt.resize((size_t)N * ldct))
        return c->fail("out of device memory");
    std::vector<double> hc((size_t)M * ldc, 0.0);
    {
        const double nan = std::nan("");
        std::vector<double> pa((size_t)M * ldk, nan), pb((size_t)N * ldk, nan);
        for (int i = 0; i < M; ++i)
            std::memcpy(pa.data() + (size_t)i * ldk, A + (size_t)i * K, (size_t)K * sizeof(double));
        for (int i = 0; i < N; ++i)
            std::memcpy(pb.data() + (size_t)i * ldk, B + (size_t)i * K, (size_t)K * sizeof(double));
        for (int i = 0; i < M; ++i)   // C on entry: what GEMM_ACCUM adds to and what GEMM_LOWER leaves alone above the diagonal
            std::memcpy(hc.data() + (size_t)i * ldc, C + (size_t)i * N, (size_t)N * sizeof(double));
        dev::h2d(dA.p, pa.data(), dA.bytes());
        dev::h2d(dB.p, pb.data(), dB.bytes());
        dev::h2d(dC.p, hc.data(), dC.bytes());
        std::string e0 = dev::sync();
        if (!e0.empty())
            return c->fail(e0);
    }
    dev::zero(dCt.p, dCt.bytes());
    GemmOp op{};
    op.A = dA.p;
    op.B = dB.p;
    op.C = dC.p;
    op.Ct = dCt.p;
    op.ldct = ldct;
    op.lda = ldk;
    op.ldb = ldk;
    op.ldc = ldc;
    op.M = M;
    op.N = N;
    op.K = K;
    op.flags = flags;
    const int T = tile_dim(shape);
    op.tiles_m = (M + T - 1) / T;
    op.tiles_n = (N + T - 1) / T;
    std::vector<GemmTile> tl;
    for (int tm = 0; tm < op.tiles_m; ++tm)
        for (int tn = 0; tn < op.tiles_n; ++tn)
            if (!((flags & GEMM_LOWER) && tm * T + (T - 1) < tn * T))
                tl.push_back(GemmTile{0, (uint16_t)tm, (uint16_t)tn});
    if (!dtl.upload(tl))
        return c->fail("out of device memory");
    if (!dev::encode_tma_2d(&op.tmA, op.A, M, K, ldk, T) || !dev::encode_tma_2d(&op.tmB, op.B, N, K, ldk, T))
        return c->fail("tensor-map encoding failed");
    dev::h2d(dop.p, &op, sizeof(op));
    launch_gemm(dop.p, 1, dtl.p, (int)tl.size(), shape, dev::stream());
    dev::d2h(hc.data(), dC.p, dC.bytes());
    std::vector<double> hct;
    if (flags & GEMM_DUAL) {
        hct.resize((size_t)N * ldct);
        dev::d2h(hct.data(), dCt.p, dCt.bytes());
    }
    std::string e = dev::sync();
    if (!e.empty())
        return c->fail(e);
    if (reps > 0) {   // timing (the accumulating form keeps adding into the device copy: the results above are from call 1)
        dev::event_record(c->ev[0]);
        for (int i = 0; i < reps; ++i)
            launch_gemm(dop.p, 1, dtl.p, (int)tl.size(), shape, dev::stream());
        dev::event_record(c->ev[1]);
        e = dev::sync();
        if (!e.empty())
            return c->fail(e);
        if (ms)
            *ms = dev::event_elapsed_ms(c->ev[0], c->ev[1]) / (float)reps;
    }
    for (int i = 0; i < M; ++i)
        std::memcpy(C + (size_t)i * N, hc.data() + (size_t)i * ldc, (size_t)N * sizeof(double));
    if (flags & GEMM_DUAL)
        for (int j = 0; j < N; ++j)
            std::memcpy(Ct + (size_t)j * M, hct.data() + (size_t)j * ldct, (size_t)M * sizeof(double));
    return 0;
}

int gadj_test_gemm(gadj_ctx* c, const double* A, const double* B, double* C, int M, int N, int K, int reps, float* ms)
{
    if (M > 0 && N > 0)
        std::memset(C, 0, (size_t)M * N * sizeof(double));
    return gadj_test_gemm_ex(c, A, B, C, nullptr, M, N, K, 0, 128, reps < 1 ? 1 : reps, ms);
}

}  // extern "C"
