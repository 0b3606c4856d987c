// tests/hostsim/kernels_host.cpp — TEST INFRASTRUCTURE.
// Plain-loop stand-ins for every launch in csrc/kernels.h: an executable statement of what
// each CUDA kernel must compute, used to validate the planner / index maps / control flow on
// CPU.  Never linked into the product library.
#include <chrono>
#include <cmath>
#include <cstring>
#include <thread>
#include <vector>

#include "../../dynadjust_b200/csrc/geodesy.h"
#include "../../dynadjust_b200/csrc/kernels.h"
#include "../../dynadjust_b200/csrc/rows.h"

namespace gadj {

void launch_gemm(const GemmOp* ops, int nops, const GemmTile* tiles, int ntiles, int shape, void*)
{
    const int TILE_M = tile_dim(shape), TILE_N = tile_dim(shape);   // the launch's tile shape (shadows the 128 defaults)
    // tile by tile, exactly the work list the persistent CTAs stride through: a tile missing from the planner's
    // list, or listed twice, changes the results
    std::vector<double> out((size_t)TILE_M * TILE_N);
    for (int t = 0; t < ntiles; ++t) {
        const GemmTile& tl = tiles[t];
        if (tl.op < 0 || tl.op >= nops)
            continue;
        const GemmOp& op = ops[tl.op];
        const int row0 = tl.tm * TILE_M, col0 = tl.tn * TILE_N;
        // the kernel's per-tile K range (triangular operands): nothing outside it may be read
        int k_lo = 0, k_hi = op.K;
        if (op.flags & GEMM_KLO_ROW)
            k_lo = row0;
        if (op.flags & GEMM_KLO_MAX)
            k_lo = row0 > col0 ? row0 : col0;
        if (op.flags & GEMM_KHI_ROW)
            k_hi = row0 + TILE_M < op.K ? row0 + TILE_M : op.K;
        k_lo = (k_lo / TILE_K) * TILE_K;
        const int i1 = row0 + TILE_M < op.M ? row0 + TILE_M : op.M, j1 = col0 + TILE_N < op.N ? col0 + TILE_N : op.N;
        for (int i = row0; i < i1; ++i)
            for (int j = col0; j < j1; ++j) {
                double acc = 0.0;
                const double* a = op.A + (int64_t)i * op.lda;
                const double* b = op.B + (int64_t)j * op.ldb;
                for (int k = k_lo; k < k_hi; ++k)
                    acc += a[k] * b[k];
                out[(size_t)(i - row0) * TILE_N + (j - col0)] = (op.flags & GEMM_NEG) ? -acc : acc;
            }
        for (int i = row0; i < i1; ++i)
            for (int j = col0; j < j1; ++j) {
                if ((op.flags & GEMM_LOWER) && i + op.tri_off < j)
                    continue;
                const double v = out[(size_t)(i - row0) * TILE_N + (j - col0)];
                if (op.flags & GEMM_SCATTER) {
                    const ScatterTarget& tg = op.tgt[op.coltgt[j / 3]];
                    int64_t r = 3ll * tg.rowmap[i / 3 - tg.jb] + i % 3;
                    int64_t c = 3ll * tg.rowmap[j / 3 - tg.jb] + j % 3;
                    tg.C[r * tg.ldc + c] += v;
                } else if (op.flags & GEMM_ACCUM)
                    op.C[(int64_t)i * op.ldc + j] += v;
                else {
                    op.C[(int64_t)i * op.ldc + j] = v;
                    if (op.flags & GEMM_DUAL)
                        op.Ct[(int64_t)j * op.ldct + i] = v;
                }
            }
    }
}

void launch_diag(const DiagOp* ops, int nops, int* info, void*)
{
    for (int o = 0; o < nops; ++o) {
        const DiagOp& op = ops[o];
        const int w = op.w;
        double* D = op.D;
        const int64_t ld = op.ldd;
        if (op.factor) {
            for (int j = 0; j < w; ++j) {
                double d = D[j * ld + j];
                for (int k = 0; k < j; ++k)
                    d -= D[j * ld + k] * D[j * ld + k];
                if (!(d > 0.0)) {
                    if (info[0] == 0)
                        info[0] = op.front + 1;
                    d = 1.0;
                }
                d = std::sqrt(d);
                D[j * ld + j] = d;
                for (int i = j + 1; i < w; ++i) {
                    double s = D[i * ld + j];
                    for (int k = 0; k < j; ++k)
                        s -= D[i * ld + k] * D[j * ld + k];
                    D[i * ld + j] = s / d;
                }
            }
        }
        if (op.W || op.Wt) {
            std::vector<double> W((size_t)w * w, 0.0);
            for (int j = 0; j < w; ++j) {
                W[(size_t)j * w + j] = 1.0 / D[j * ld + j];
                for (int i = j + 1; i < w; ++i) {
                    double s = 0.0;
                    for (int k = j; k < i; ++k)
                        s += D[i * ld + k] * W[(size_t)k * w + j];
                    W[(size_t)i * w + j] = -s / D[i * ld + i];
                }
            }
            for (int i = 0; i < w; ++i)
                for (int j = 0; j < w; ++j) {
                    if (op.W)
                        op.W[i * op.ldw + j] = W[(size_t)i * w + j];
                    if (op.Wt)
                        op.Wt[j * op.ldwt + i] = W[(size_t)i * w + j];
                }
        }
    }
}

void launch_push(const PushOp* ops, int nops, int, const PeerTable* pt, double* const* bases, void*)
{
    for (int o = 0; o < nops; ++o) {
        const PushOp& op = ops[o];
        double* base = bases[op.buf] + op.off;
        for (int r = 0; r < op.rows; ++r)
            for (int c = 0; c < op.cols; ++c) {
                double* p = base + (int64_t)r * op.ld + c;
                for (int q = 0; q < pt->nranks; ++q)
                    if (q != pt->rank && q != op.skip && (op.target < 0 || q == op.target))
                        *reinterpret_cast<double*>(reinterpret_cast<char*>(p) + pt->delta[op.buf][q]) = *p;
            }
    }
}

// ---- multi-rank stand-ins: the ranks are threads of one process or separate processes sharing memory ----
void launch_barrier(const PeerTable* pt, unsigned long long target, int* info, void*)
{
    for (int q = 0; q < pt->nranks; ++q)
        __atomic_fetch_add(pt->counter[q], 1ull, __ATOMIC_SEQ_CST);
    const auto t0 = std::chrono::steady_clock::now();
    while (__atomic_load_n(pt->counter[pt->rank], __ATOMIC_SEQ_CST) < target) {
        std::this_thread::yield();
        if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(60)) {
            info[1] = 1;
            break;
        }
    }
}

void launch_allreduce(const ReduceOp* ops, int nops, const PeerTable* pt, double* base, int buf, void*)
{
    const int n = pt->nranks, me = pt->rank;
    for (int o = 0; o < nops; ++o) {
        const uint64_t lo = ops[o].count * me / n, hi = ops[o].count * (me + 1) / n;
        for (uint64_t i = lo; i < hi; ++i) {
            double* p = base + ops[o].off + i;
            double s = 0.0;
            for (int q = 0; q < n; ++q)
                s += *reinterpret_cast<const double*>(reinterpret_cast<const char*>(p) + pt->delta[buf][q]);
            for (int q = 0; q < n; ++q)
                *reinterpret_cast<double*>(reinterpret_cast<char*>(p) + pt->delta[buf][q]) = s;
        }
    }
}

void launch_share_info(const PeerTable* pt, int* info, void*)
{
    for (int q = 0; q < pt->nranks; ++q)
        if (q != pt->rank)
            for (int k = 0; k < 2; ++k)
                if (info[k] != 0) {
                    int cur = __atomic_load_n(pt->info[q] + k, __ATOMIC_SEQ_CST);
                    while (cur < info[k] && !__atomic_compare_exchange_n(pt->info[q] + k, &cur, info[k], false, __ATOMIC_SEQ_CST,
                                                                         __ATOMIC_SEQ_CST)) {
                    }
                }
}

void launch_trimv(const TrimvOp* ops, int nops, void*)
{
    for (int o = 0; o < nops; ++o) {
        const TrimvOp& op = ops[o];
        for (int i = 0; i < op.nrows; ++i) {
            const int row = op.row0 + i;
            const int c_lo = op.upper ? row : 0, c_hi = op.upper ? op.k : row + 1;
            double s = 0.0;
            for (int c = c_lo; c < c_hi; ++c)
                s += op.A[(int64_t)i * op.ld + c] * op.x[c];
            op.y[row] = s;
        }
    }
}

void launch_gemv(const GemvOp* ops, int nops, const double* x_ro, double* x, int backward, void*)
{
    for (int o = 0; o < nops; ++o) {
        const GemvOp& op = ops[o];
        if (!backward) {
            for (int i = 0; i < op.nrows; ++i) {
                double s = 0.0;
                for (int c = 0; c < op.w; ++c)
                    s += op.P[(int64_t)i * op.ld + c] * op.xj[c];
                x[op.rowidx[i]] -= s;
            }
        } else {
            for (int c = 0; c < op.w; ++c) {
                double s = 0.0;
                for (int i = 0; i < op.nrows; ++i)
                    s += op.P[(int64_t)i * op.ld + c] * x_ro[op.rowidx[i]];
                op.xj[c] -= s;
            }
        }
    }
}

void launch_transpose(const TransposeOp* ops, int nops, int, void*)
{
    for (int o = 0; o < nops; ++o) {
        const TransposeOp& op = ops[o];
        for (int r = 0; r < op.rows; ++r)
            for (int c = 0; c < op.cols; ++c)
                op.dst[(int64_t)c * op.ldd + r] = op.src[(int64_t)r * op.lds + c];
    }
}

void launch_gather(const GatherOp* ops, int nops, const GatherTile* tiles, int ntiles, void*)
{
    // tile by tile, exactly the planner's list (a missing tile leaves part of G unwritten and shows in the results)
    const int GT = GATHER_TILE_STATIONS;
    for (int t = 0; t < ntiles; ++t) {
        if (tiles[t].op < 0 || tiles[t].op >= nops)
            continue;
        const GatherOp& op = ops[tiles[t].op];
        const int i_lo = op.jb + tiles[t].ti * GT, j_lo = op.jb + tiles[t].tj * GT;
        for (int i = i_lo; i < i_lo + GT && i < op.nb; ++i) {
            const int64_t zr = 3ll * op.rowmap[i - op.jb];
            for (int j = j_lo; j < j_lo + GT && j < op.je && j <= i; ++j) {
                const int64_t zc = 3ll * op.rowmap[j - op.jb];
                for (int a = 0; a < 3; ++a)
                    for (int b = 0; b < 3; ++b) {
                        double v;
                        if (i == j && a < b)
                            v = op.Z[(zr + b) * op.ld + zc + a];
                        else
                            v = op.Z[(zr + a) * op.ld + zc + b];
                        op.G[(3ll * i + a) * op.ldg + 3 * j + b] = v;
                        op.G[(3ll * j + b) * op.ldg + 3 * i + a] = v;
                    }
            }
        }
    }
}

void launch_init_normals(const double* cblock, double* ndiag, double* noff, double* w, uint32_t nstn, uint64_t nedge, void*)
{
    if (cblock) {
        std::memcpy(ndiag, cblock, 9 * (size_t)nstn * sizeof(double));
        std::memset(noff, 0, 9 * (size_t)nedge * sizeof(double));
    }
    std::memset(w, 0, 3 * (size_t)nstn * sizeof(double));
}

void launch_assemble_g(const AssembleParams& p, void*)
{
    // pass 1: per baseline
    for (uint64_t b = 0; b < p.nbaselines; ++b) {
        const dna_msr_t* m = p.msr + p.first[b];
        const uint32_t s1 = m[0].station1, s2 = m[0].station2;
        double l[3];
        for (int r = 0; r < 3; ++r)
            l[r] = m[r].term1 - (p.est[3 * (size_t)s2 + r] - p.est[3 * (size_t)s1 + r]);
        const double up[6] = {m[0].term2, m[1].term2, m[2].term2, m[1].term3, m[2].term3, m[2].term4};
        double q[6];
        if (!spd3_inverse(up, q))
            for (double& v : q)
                v = NAN;
        const double V[9] = {q[0], q[1], q[2], q[1], q[3], q[4], q[2], q[4], q[5]};
        double* slot = p.bq + 9 * b;
        for (int r = 0; r < 3; ++r)
            slot[6 + r] = V[3 * r] * l[0] + V[3 * r + 1] * l[1] + V[3 * r + 2] * l[2];
        if (p.normals) {
            for (int k = 0; k < 6; ++k)
                slot[k] = q[k];
            const uint32_t ew = p.edge[b];
            if (!(ew & EDGE_EXCLUSIVE)) {
                double* o = p.noff + 9 * (size_t)(ew & EDGE_SLOT_MASK);
                for (int k = 0; k < 9; ++k)
                    o[k] -= V[k];
            }
        }
    }
}

void launch_station_sum(const AssembleParams& p, void*)
{
    if (p.nbaselines == 0)
        return;
    // pass 2: per station, over its incidence list
    static const int sym[9] = {0, 1, 2, 1, 3, 4, 2, 4, 5};
    for (uint32_t s = 0; s < p.nstn; ++s)
        for (uint32_t i = p.inc_ptr[s]; i < p.inc_ptr[s + 1]; ++i) {
            const uint32_t e = p.inc[i];
            const double* slot = p.bq + 9ull * (e & 0x7FFFFFFFu);
            if (p.normals)
                for (int k = 0; k < 9; ++k)
                    p.ndiag[9 * (size_t)s + k] += slot[sym[k]];
            for (int r = 0; r < 3; ++r)
                p.w[3 * (size_t)s + r] += (e & 0x80000000u) ? slot[6 + r] : -slot[6 + r];
        }
}

// rows / clusters: the arithmetic is shared with the kernels (rows.h); only the thread loops differ
void launch_rows(const RowsParams& p, void*)
{
    const Ellipsoid el = make_ellipsoid(p.semi_major, p.inv_flattening);
    for (uint64_t i = 0; i < p.nrows; ++i)
        row_body(p, i, el);
}

void launch_rows_stats(const RowsParams& p, void*)
{
    const Ellipsoid el = make_ellipsoid(p.semi_major, p.inv_flattening);
    double acc[4] = {0, 0, 0, 0};
    for (uint64_t i = 0; i < p.nrows; ++i)
        row_stats_body(p, i, el, acc);
    for (int k = 0; k < 4; ++k)
        p.sums[k] += acc[k];
}

void launch_clusters(const ClusterParams& p, void*)
{
    for (uint32_t ci = 0; ci < p.nclusters; ++ci) {
        const ClusterDesc c = p.clusters[ci];
        for (uint32_t r = 0; r < c.n; ++r)
            cluster_t_body(p, c, r);
        for (uint32_t j = 0; j < c.ns; ++j)
            cluster_rhs_body(p, c, j);
        if (p.normals)
            for (uint64_t q = 0; q < (uint64_t)c.ns * (c.ns + 1) / 2; ++q)
                cluster_pair_body(p, c, q);
    }
}

void launch_cluster_chi(const ClusterParams& p, void*)
{
    for (uint32_t ci = 0; ci < p.nclusters; ++ci) {
        const ClusterDesc c = p.clusters[ci];
        if (c.type == 'D')
            continue;
        for (uint32_t r = 0; r < c.n; ++r) {
            cluster_t_body(p, c, r);
            p.sums[0] += p.row_l[c.row0 + r] * p.row_t[c.row0 + r];
        }
    }
}

// plain Cholesky inverse (factor, invert the factor, W^T W) per cluster matrix
void launch_cluster_inverse(const ClusterDesc* clusters, uint32_t nclusters, double* work, double* out, int* info, void*)
{
    for (uint32_t ci = 0; ci < nclusters; ++ci) {
        const uint32_t n = clusters[ci].n;
        double* A = work + clusters[ci].vinv_off;
        double* O = out + clusters[ci].vinv_off;
        std::vector<double> L((size_t)n * n, 0.0), W((size_t)n * n, 0.0);
        bool bad = false;
        for (uint32_t j = 0; j < n; ++j) {
            double d = A[(size_t)j * n + j];
            for (uint32_t k = 0; k < j; ++k)
                d -= L[(size_t)j * n + k] * L[(size_t)j * n + k];
            if (!(d > 0.0)) {
                bad = true;
                d = 1.0;
            }
            L[(size_t)j * n + j] = std::sqrt(d);
            for (uint32_t i = j + 1; i < n; ++i) {
                double s = A[(size_t)i * n + j];
                for (uint32_t k = 0; k < j; ++k)
                    s -= L[(size_t)i * n + k] * L[(size_t)j * n + k];
                L[(size_t)i * n + j] = s / L[(size_t)j * n + j];
            }
        }
        for (uint32_t c = 0; c < n; ++c) {
            W[(size_t)c * n + c] = 1.0 / L[(size_t)c * n + c];
            for (uint32_t i = c + 1; i < n; ++i) {
                double s = 0.0;
                for (uint32_t k = c; k < i; ++k)
                    s += L[(size_t)i * n + k] * W[(size_t)k * n + c];
                W[(size_t)i * n + c] = -s / L[(size_t)i * n + i];
            }
        }
        for (uint32_t i = 0; i < n; ++i)
            for (uint32_t j = 0; j <= i; ++j) {
                double s = 0.0;
                for (uint32_t k = i; k < n; ++k)
                    s += W[(size_t)k * n + i] * W[(size_t)k * n + j];
                O[(size_t)i * n + j] = O[(size_t)j * n + i] = s;
            }
        if (bad && info[0] == 0)
            info[0] = (int)ci + 1;
    }
}

void launch_compute_scale(const ScatterParams& p, void*)
{
    for (uint32_t s = 0; s < p.nstn; ++s)
        for (int c = 0; c < 3; ++c)
            p.dscale[3 * (size_t)s + c] = p.scale ? 1.0 / std::sqrt(p.ndiag[9 * (size_t)s + 4 * c]) : 1.0;
}

void launch_scatter_normals(const ScatterParams& p, void*)
{
    for (uint32_t s = 0; s < p.nstn; ++s) {
        if (p.diag_dest[s] == ~0ull)
            continue;
        double* d = p.panels + p.diag_dest[s];
        const uint32_t ld = p.diag_ld[s];
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c)
                d[(size_t)r * ld + c] = p.ndiag[9 * (size_t)s + 3 * r + c] * p.dscale[3 * (size_t)s + r] * p.dscale[3 * (size_t)s + c];
    }
    for (uint64_t e = 0; e < p.nedge; ++e) {
        if (p.off_dest[e] == ~0ull)
            continue;
        double* d = p.panels + p.off_dest[e];
        const uint32_t ld = p.off_ld[e];
        const uint32_t hi = p.edge_hi[e], lo = p.edge_lo[e];
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c)
            {
                const uint32_t b = p.edge_bsl[e];
                const int k = 3 * r + c;
                const double v = b != ~0u ? -p.bq[9ull * b + GADJ_SYM3(k)] : p.noff[9 * e + k];
                d[(size_t)r * ld + c] = v * p.dscale[3 * (size_t)hi + r] * p.dscale[3 * (size_t)lo + c];
            }
    }
}

void launch_permute_rhs(const double* w, const double* dscale, const uint32_t* pos, const uint8_t* pos_owned, double* b,
                        uint32_t nstn, void*)
{
    for (uint32_t s = 0; s < nstn; ++s)
        for (int c = 0; c < 3; ++c)
            b[3 * (size_t)pos[s] + c] = (!pos_owned || pos_owned[pos[s]]) ? dscale[3 * (size_t)s + c] * w[3 * (size_t)s + c] : 0.0;
}

void launch_mask_positions(double* x, const uint8_t* pos_owned, uint32_t nstn, void*)
{
    for (size_t i = 0; i < 3 * (size_t)nstn; ++i)
        if (!pos_owned[i / 3])
            x[i] = 0.0;
}

void launch_apply_corrections(const double* x, const double* dscale, const uint32_t* pos, double* corr, double* est,
                              uint32_t nstn, double*, void*)
{
    size_t best = 0;
    for (uint32_t s = 0; s < nstn; ++s)
        for (int c = 0; c < 3; ++c) {
            size_t i = 3 * (size_t)s + c;
            corr[i] = dscale[i] * x[3 * (size_t)pos[s] + c];
            est[i] += corr[i];
            if (std::fabs(corr[i]) > std::fabs(corr[best]))
                best = i;
        }
    corr[3 * (size_t)nstn] = corr[best];
    corr[3 * (size_t)nstn + 1] = (double)best;
    for (int c = 0; c < 3; ++c)
        corr[3 * (size_t)nstn + 2 + c] = corr[(best / 3) * 3 + c];
}

void launch_extract_station_vcv(const double* panels, const uint64_t* diag_dest, const uint32_t* diag_ld,
                                const double* dscale, double* vcv, uint32_t nstn, void*)
{
    for (uint32_t s = 0; s < nstn; ++s) {
        if (diag_dest[s] == ~0ull) {
            for (int k = 0; k < 9; ++k)
                vcv[9 * (size_t)s + k] = 0.0;
            continue;
        }
        const double* z = panels + diag_dest[s];
        const uint32_t ld = diag_ld[s];
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) {
                int a = r > c ? r : c, b = r > c ? c : r;
                vcv[9 * (size_t)s + 3 * r + c] = z[(size_t)a * ld + b] * dscale[3 * (size_t)s + r] * dscale[3 * (size_t)s + c];
            }
    }
}

void launch_extract_edge_vcv(const double* panels, const uint64_t* off_dest, const uint32_t* off_ld, const uint32_t* edge_hi,
                             const uint32_t* edge_lo, const double* dscale, double* q, uint64_t nedge, void*)
{
    for (uint64_t e = 0; e < nedge; ++e) {
        if (off_dest[e] == ~0ull) {
            for (int k = 0; k < 9; ++k)
                q[9 * e + k] = 0.0;
            continue;
        }
        const double* z = panels + off_dest[e];
        const uint32_t ld = off_ld[e];
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c)
                q[9 * e + 3 * r + c] = z[(size_t)r * ld + c] * dscale[3 * (size_t)edge_hi[e] + r] * dscale[3 * (size_t)edge_lo[e] + c];
    }
}

void launch_stats_g(const StatsParams& p, void*)
{
    for (uint64_t b = 0; b < p.nbaselines; ++b) {
        dna_msr_t* m = p.msr + p.first[b];
        const uint32_t s1 = m[0].station1, s2 = m[0].station2;
        double l[3];
        for (int r = 0; r < 3; ++r)
            l[r] = m[r].term1 - (p.est[3 * (size_t)s2 + r] - p.est[3 * (size_t)s1 + r]);
        const uint32_t ew = p.edge[b];
        const double* Qo = p.vcv_off + 9 * (size_t)(ew & EDGE_SLOT_MASK);
        const bool s1_is_hi = (ew & 0x80000000u) != 0;
        const double* Q11 = p.vcv_diag + 9 * (size_t)s1;
        const double* Q22 = p.vcv_diag + 9 * (size_t)s2;
        // Q21(i,j) = N^-1[s2+i, s1+j]
        auto Q21 = [&](int i, int j) { return s1_is_hi ? Qo[3 * j + i] : Qo[3 * i + j]; };
        // Precision_Adjusted_GNSS_bsl (dnatemplatematrixfuncs.hpp:255-297): A Q A^T, A = [-I I]
        double P[3][3];
        for (int i = 0; i < 3; ++i)
            for (int j = i; j < 3; ++j) {
                double t0 = -Q11[3 * i + j] + Q21(i, j);       // tmp[i][j]
                double t1 = -Q21(j, i) + Q22[3 * i + j];       // tmp[i][j+3]: -Q(s1+i, s2+j) + Q(s2+i, s2+j)
                P[i][j] = t1 - t0;
            }
        const double prec[3] = {P[0][0], P[1][1], P[2][2]};
        const double mprec[3] = {m[0].term2, m[1].term3, m[2].term4};
        for (int r = 0; r < 3; ++r) {
            // UpdateMsrRecord / UpdateMsrRecordStats (ADJ:8187-8298)
            m[r].measCorr = -l[r];
            m[r].measAdj = m[r].term1 + m[r].measCorr;
            m[r].measAdjPrec = prec[r];
            double rp = mprec[r] - prec[r];
            if (rp < 0.0)
                rp = std::fabs(rp);
            m[r].residualPrec = rp;
            double pel = std::sqrt(mprec[r]) / std::sqrt(rp);
            if (pel < 0. || pel > 700.)
                pel = 999.99;
            m[r].NStat = m[r].measCorr / std::sqrt(rp);
            if (std::fabs(m[r].NStat) > p.critical)
                p.sums[3] += 1.0;
            // ComputeGlobalPelzer_GXY (ADJ:8396-8427)
            if (pel > 0. && pel < 999.99) {
                p.sums[1] += pel * pel - 1.;
                p.sums[2] += 1.0;
            } else
                pel = 999.99;
            m[r].PelzerRel = pel;
        }
        const double up[6] = {m[0].term2, m[1].term2, m[2].term2, m[1].term3, m[2].term3, m[2].term4};
        double q[6];
        if (!spd3_inverse(up, q))
            for (double& v : q)
                v = NAN;
        const double V[9] = {q[0], q[1], q[2], q[1], q[3], q[4], q[2], q[4], q[5]};
        double cs = 0.0;
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c)
                cs += V[3 * r + c] * l[r] * l[c];
        p.sums[0] += cs;
    }
}

void launch_cart_to_geo(const double* est, double* llh, uint32_t nstn, double a, double invf, void*)
{
    Ellipsoid e = make_ellipsoid(a, invf);
    for (uint32_t s = 0; s < nstn; ++s)
        cart_to_geo(e, est[3 * (size_t)s], est[3 * (size_t)s + 1], est[3 * (size_t)s + 2], llh + 3 * (size_t)s);
}

}  // namespace gadj
