// plan.cpp — static launch lists for factorisation, triangular solves and the
// selected inverse on the supernode tree.
//
// Numerically this is the block form of what the reference does per block with
// dense LAPACK calls: Solve()'s dpotrf (ADJ:6628, MATC:982) becomes the blocked
// right-looking Cholesky of each front's pivot block; the junction carry
// N_next[JSL,JSL] += J^-1 (ADJ:1076-1126) becomes the Schur update
// -L21 L21^T scattered into the ancestors' panels; dpotri (MATC:984) becomes the
// top-down selected inverse Z11 = W^T W + Y^T Z22 Y, Z21 = -Z22 Y with
// W = L11^-1, Y = L21 W  (the reverse/combine passes, ADJ:3461-3590).
// W (and its transpose) of every front is formed once, right after the front's pivot block is factorised, and
// kept: the substitutions multiply by it (two launches per tree level and direction instead of two per
// pivot tile) and the selected inverse starts from it.
#include "plan.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>

#include "dev.h"

namespace gadj {
namespace {

inline size_t al16(size_t n) { return (n + 15) & ~(size_t)15; }
inline uint32_t even(uint32_t n) { return n + (n & 1u); }
inline int cdiv(int a, int b) { return (a + b - 1) / b; }

struct SelinvWs {
    size_t G, Yt, Z21t, total;
    uint32_t ldg, ldr;
};

SelinvWs ws_layout(const Front& f)
{
    SelinvWs w{};
    w.ldg = even(f.r);
    w.ldr = even(f.r);
    size_t o = 0;
    if (f.r > 0) {
        w.G = o;
        o += al16((size_t)f.r * w.ldg);
        w.Yt = o;
        o += al16((size_t)f.k * w.ldr);
        w.Z21t = o;
        o += al16((size_t)f.k * w.ldr);
    }
    w.total = o;
    return w;
}

// persistent inverse pivot blocks: W then Wt, each k x ldw row-major
inline uint32_t ldw_of(const Front& f) { return even(f.k); }
inline size_t wblock(const Front& f) { return al16((size_t)f.k * ldw_of(f)); }

struct Builder {
    const Symbolic& s;
    const PlanBuffers& b;
    Plan& p;
    std::string err;

    std::vector<size_t> woff;   // per front: offset of its W block in b.wbuf (Wt follows at + wblock)

    // per op of a batch: how its tiles are shared out among the ranks (multi-GPU, replicated top fronts)
    enum Dist { D_ALL = 0, D_SUM = 1 };
    struct Share {
        int dist;   // D_ALL: this rank runs every tile; D_SUM: tile (tm, tn) goes to rank (tm + tn) % world
        int push;   // the finished tiles are pushed into every peer's replica: McBuf of C in bits 0..7, of Ct in bits 8..15
    };
    std::vector<Share> share;   // parallel to the pending GEMM batch
    std::vector<PushOp> pushes; // pending pushes (the tiles of the launches since the last flush_push)

    Builder(const Symbolic& S, const PlanBuffers& B, Plan& P) : s(S), b(B), p(P)
    {
        if (const char* e = getenv("GADJ_TILE_LONG_K"))   // tuning aid: K (in 16-deep steps) from which 128 x 128 tiles are kept
            tile_long_ksteps = atof(e);
        // top fronts first: the same offsets on every rank (see finalize_layout)
        woff.assign(s.fronts.size(), 0);
        size_t o = 0;
        for (int pass = 0; pass < 2; ++pass)
            for (size_t f = 0; f < s.fronts.size(); ++f) {
                const Front& fr = s.fronts[f];
                if ((pass == 0) != (fr.top != 0))
                    continue;
                if (fr.top || fr.owner == s.rank) {
                    woff[f] = o;
                    o += 2 * wblock(fr);
                }
            }
    }

    double* panel(const Front& f) const { return b.panels + f.panel_off; }
    double* Wof(uint32_t fi) const { return b.wbuf + woff[fi]; }
    double* Wtof(uint32_t fi) const { return b.wbuf + woff[fi] + wblock(s.fronts[fi]); }

    // fronts of a level inside this rank's subtrees / the replicated top fronts of a level (every rank works on those)
    std::vector<uint32_t> local_fronts(size_t lv) const
    {
        std::vector<uint32_t> v;
        for (uint32_t f : s.levels[lv])
            if (s.fronts[f].owner == s.rank && !s.fronts[f].top)
                v.push_back(f);
        return v;
    }
    std::vector<uint32_t> top_fronts(size_t lv) const
    {
        std::vector<uint32_t> v;
        for (uint32_t f : s.levels[lv])
            if (s.fronts[f].top)
                v.push_back(f);
        return v;
    }
    bool multi() const { return s.world > 1; }

    void add_barrier(std::vector<Launch>& out, int level)
    {
        if (!multi())
            return;
        Launch L{};
        L.kind = L_BARRIER;
        L.level = level;
        out.push_back(L);
        p.barriers++;
    }
    // barrier; sum over the ranks of the given ranges of a replicated buffer; barrier
    void add_allreduce(std::vector<Launch>& out, int level, int buf, const std::vector<ReduceOp>& ranges)
    {
        if (!multi() || ranges.empty())
            return;
        add_barrier(out, level);
        Launch L{};
        L.kind = L_ALLREDUCE;
        L.level = level;
        L.buf = buf;
        L.op_begin = (int64_t)p.reduce.size();
        L.op_count = (int32_t)ranges.size();
        p.reduce.insert(p.reduce.end(), ranges.begin(), ranges.end());
        for (const ReduceOp& r : ranges) {   // this rank's slice: read from and stored into world - 1 peers
            p.nvlink_read_bytes += 8.0 * (double)r.count * (s.world - 1) / s.world;
            p.nvlink_write_bytes += 8.0 * (double)r.count * (s.world - 1) / s.world;
        }
        out.push_back(L);
        add_barrier(out, level);
    }

    double* buffer_base(int buf) const { return buf == MC_PANELS ? b.panels : buf == MC_WBUF ? b.wbuf : buf == MC_POOL ? b.pool : nullptr; }
    void add_push(int buf, const double* ptr, int64_t ld, int rows, int cols, int target = -1, int skip = -1)
    {
        if (!multi() || rows <= 0 || cols <= 0)
            return;
        PushOp o{};
        o.off = (uint64_t)(ptr - buffer_base(buf));
        o.ld = ld;
        o.rows = rows;
        o.cols = cols;
        o.buf = buf;
        o.target = (int16_t)target;
        o.skip = (int16_t)skip;
        pushes.push_back(o);
        const int ndst = target >= 0 ? 1 : s.world - 1 - (skip >= 0 ? 1 : 0);
        p.nvlink_write_bytes += 8.0 * rows * cols * ndst;
    }
    // Broadcast of a block its owner has finished, in two launches: the owner sends each of the other ranks one slice of
    // the rows (its NVLink egress carries the block once), then every rank forwards the slice it received to the others.
    // stage 0 is planned by the owner, stage 1 by everybody else; a barrier (flush_push) follows each stage.
    void add_broadcast(int stage, int owner, int buf, const double* ptr, int64_t ld, int rows, int cols)
    {
        const int P = s.world;
        auto slice = [&](int q, int& r0, int& r1) {
            const int idx = q < owner ? q : q - 1;   // the ranks other than the owner, in order
            r0 = (int)((int64_t)rows * idx / (P - 1));
            r1 = (int)((int64_t)rows * (idx + 1) / (P - 1));
        };
        int r0, r1;
        if (stage == 0 && s.rank == owner) {
            for (int q = 0; q < P; ++q) {
                if (q == owner)
                    continue;
                slice(q, r0, r1);
                add_push(buf, ptr + (int64_t)r0 * ld, ld, r1 - r0, cols, q);
            }
        } else if (stage == 1 && s.rank != owner && P > 2) {
            slice(s.rank, r0, r1);
            add_push(buf, ptr + (int64_t)r0 * ld, ld, r1 - r0, cols, -1, owner);
        }
    }
    // the pending pushes as one launch, then a barrier: afterwards every replica holds what the ranks have just finished.
    // Every rank calls this at the same points of the plan (a rank with nothing to push still meets the others).
    void flush_push(std::vector<Launch>& out, int level)
    {
        if (!multi())
            return;
        if (!pushes.empty()) {
            Launch L{};
            L.kind = L_PUSH;
            L.level = level;
            L.op_begin = (int64_t)p.push.size();
            L.op_count = (int32_t)pushes.size();
            int maxrows = 0;
            for (const PushOp& o : pushes)
                maxrows = std::max(maxrows, o.rows);
            L.total_tiles = std::max(1, std::min(64, maxrows / 16));   // CTAs per op
            p.push.insert(p.push.end(), pushes.begin(), pushes.end());
            pushes.clear();
            out.push_back(L);
        }
        add_barrier(out, level);
    }

    void add_gemm(std::vector<GemmOp>& batch, const double* A, int64_t lda, const double* B, int64_t ldb, double* C,
                  int64_t ldc, int M, int N, int K, int flags, int tri_off = 0, const int32_t* coltgt = nullptr,
                  Share sh = Share{D_ALL, 0})
    {
        if (M <= 0 || N <= 0 || K <= 0)
            return;
        if (share.size() != batch.size())
            share.resize(batch.size(), Share{D_ALL, 0});
        share.push_back(sh);
        GemmOp op{};
        op.A = A;
        op.B = B;
        op.C = C;
        op.coltgt = coltgt;
        op.tgt = b.tgt;
        op.lda = lda;
        op.ldb = ldb;
        op.ldc = ldc;
        op.M = M;
        op.N = N;
        op.K = K;
        op.flags = flags;
        op.tri_off = tri_off;
        batch.push_back(op);   // tiles_m / tiles_n: flush_gemm, once the launch's tile shape is known
    }

    // K steps (16 deep) of the tile at (row0, col0) — the kernel's tile_k_range
    static int tile_ksteps(const GemmOp& op, int row0, int col0, int T)
    {
        int k_lo = 0, k_hi = op.K;
        if (op.flags & GEMM_KLO_ROW)
            k_lo = row0;
        if (op.flags & GEMM_KLO_MAX)
            k_lo = std::max(row0, col0);
        if (op.flags & GEMM_KHI_ROW)
            k_hi = std::min(row0 + T, op.K);
        return k_hi > k_lo ? cdiv(k_hi, TILE_K) - k_lo / TILE_K : 0;
    }
    static bool tile_has_work(const GemmOp& op, int tm, int tn, int T)
    {
        return !((op.flags & GEMM_LOWER) && tm * T + (T - 1) + op.tri_off < tn * T);   // else wholly above the diagonal
    }
    // Tile shape of a launch.  The tensor pipe spends the same time on a padded element as on a useful one, so the cost
    // of a launch is its tiles' area x (K steps + the epilogue's equivalent in K steps: a plain store ~2, the transposed
    // second copy ~4, the RED.ADD scatter ~12).  Measured on C4 launch by launch (profiles/r2_tile_shapes_c4.txt): the
    // 64 x 64 shape — three CTAs per SM, so one tile's epilogue runs under the others' products — wins or ties wherever
    // it does not cost more padded work, by 2x on the small fronts of the low tree levels (k ~ 40-190, r ~ 200-600, where
    // 128-wide tiles are half padding) and still by 2-15 % in the middle of the tree; only launches of long products
    // (K >= ~2000: the Z21 / Schur products of the top fronts) run 1.5-3.5 % faster on 128 x 128 tiles, whose operands
    // are read half as often.  Ops whose tiles are shared out among the ranks or pushed to the peers (top fronts) and
    // in-place products wider than 64 stay on 128 x 128.
    int choose_shape(const std::vector<GemmOp>& batch) const
    {
        for (size_t i = 0; i < batch.size(); ++i) {
            if (share[i].dist != D_ALL || share[i].push)
                return TILE_SHAPE_128;
            // in-place products (the pivot panel P <- P W^T: C is A) need every row block in ONE tile — a second column
            // tile would read rows the first one has already overwritten
            if ((batch[i].C == batch[i].A || batch[i].C == batch[i].B) && batch[i].N > tile_dim(TILE_SHAPE_64))
                return TILE_SHAPE_128;
        }
        if (b.gemm_tile == 128)
            return TILE_SHAPE_128;
        if (b.gemm_tile == 64)
            return TILE_SHAPE_64;
        double cost[2] = {0, 0}, ksteps128 = 0, tiles128 = 0;
        const size_t stride = std::max<size_t>(1, batch.size() / 512);   // a sample of the ops is enough
        for (size_t i = 0; i < batch.size(); i += stride) {
            const GemmOp& op = batch[i];
            const double epi = (op.flags & GEMM_SCATTER) ? 12.0 : (op.flags & GEMM_DUAL) ? 4.0 : 2.0;
            for (int sh = 0; sh < 2; ++sh) {
                const int T = tile_dim(sh);
                double c = 0;
                for (int tm = 0; tm < cdiv(op.M, T); ++tm)
                    for (int tn = 0; tn < cdiv(op.N, T); ++tn)
                        if (tile_has_work(op, tm, tn, T)) {
                            const int nk = tile_ksteps(op, tm * T, tn * T, T);
                            c += nk + epi;
                            if (sh == TILE_SHAPE_128) {
                                ksteps128 += nk;
                                tiles128 += 1;
                            }
                        }
                cost[sh] += c * T * T;
            }
        }
        const bool long_k = tiles128 > 0 && ksteps128 / tiles128 >= tile_long_ksteps;
        return cost[TILE_SHAPE_64] <= cost[TILE_SHAPE_128] * (long_k ? 0.97 : 1.02) ? TILE_SHAPE_64 : TILE_SHAPE_128;
    }
    double tile_long_ksteps = 128.0;   // K >= 2048

    // appends a batch as one launch; picks the tile shape, assigns tile ranges and encodes the tensor maps
    void flush_gemm(std::vector<GemmOp>& batch, std::vector<Launch>& out, int level, int tag = T_NONE)
    {
        if (batch.empty())
            return;
        Launch L{};
        L.kind = L_GEMM;
        L.tag = tag;
        L.op_begin = (int64_t)p.gemm.size();
        L.op_count = (int32_t)batch.size();
        L.level = level;
        L.tile_begin = (int64_t)p.tiles.size();
        share.resize(batch.size(), Share{D_ALL, 0});
        L.shape = choose_shape(batch);
        if (L.shape == TILE_SHAPE_64)
            p.launches_tile64++;
        const int T = tile_dim(L.shape);
        int32_t opi = 0;
        for (GemmOp& op : batch) {
            const Share sh = share[opi];
            size_t all_tiles = 0, my_tiles = 0;
            op.tiles_m = cdiv(op.M, T);
            op.tiles_n = cdiv(op.N, T);
            for (int tm = 0; tm < op.tiles_m; ++tm)
                for (int tn = 0; tn < op.tiles_n; ++tn) {
                    if (!tile_has_work(op, tm, tn, T))
                        continue;
                    ++all_tiles;
                    if (sh.dist == D_SUM && (tm + tn) % s.world != s.rank)
                        continue;   // another rank's tile
                    ++my_tiles;
                    p.tiles.push_back(GemmTile{opi, (uint16_t)tm, (uint16_t)tn});
                    if (sh.push) {
                        const int rows = std::min(T, op.M - tm * T), cols = std::min(T, op.N - tn * T);
                        const int bc = sh.push & 0xff, bt = (sh.push >> 8) & 0xff;
                        add_push(bc, op.C + (int64_t)tm * T * op.ldc + (int64_t)tn * T, op.ldc, rows, cols);
                        if (bt)
                            add_push(bt, op.Ct + (int64_t)tn * T * op.ldct + (int64_t)tm * T, op.ldct, cols, rows);
                    }
                }
            ++opi;
            // useful flops: lower-only outputs drop the strict upper triangle of the leading square;
            // triangular operands halve the K range over that square
            double M = op.M, N = op.N, K = op.K;
            double fl = 2.0 * M * N * K;
            if (op.flags & GEMM_LOWER) {
                double sq = std::min(M, N);
                fl -= sq * sq * K;
            }
            if (op.flags & (GEMM_KLO_ROW | GEMM_KHI_ROW))
                fl *= 0.5;
            if (op.flags & GEMM_KLO_MAX)
                fl = 2.0 * M * M * M / 6.0 * (op.flags & GEMM_LOWER ? 1.0 : 2.0);
            if (all_tiles)
                fl *= (double)my_tiles / (double)all_tiles;   // this rank's share of a distributed op
            L.flops += fl;
            if (!dev::encode_tma_2d(&op.tmA, op.A, (uint64_t)op.M, (uint64_t)op.K, (uint64_t)op.lda, (uint32_t)T) ||
                !dev::encode_tma_2d(&op.tmB, op.B, (uint64_t)op.N, (uint64_t)op.K, (uint64_t)op.ldb, (uint32_t)T))
                err = "tensor-map encoding failed";
            p.gemm.push_back(op);
        }
        L.total_tiles = (int32_t)(p.tiles.size() - (size_t)L.tile_begin);
        out.push_back(L);
        batch.clear();
        share.clear();
    }

    void flush_diag(std::vector<DiagOp>& batch, std::vector<Launch>& out, int level)
    {
        if (batch.empty())
            return;
        Launch L{};
        L.kind = L_DIAG;
        L.op_begin = (int64_t)p.diag.size();
        L.op_count = (int32_t)batch.size();
        L.level = level;
        for (auto& op : batch) {
            L.flops += (double)op.w * op.w * op.w * (op.factor ? 2.0 / 3.0 : 1.0 / 3.0);
            p.diag.push_back(op);
        }
        out.push_back(L);
        batch.clear();
    }

    template <class Op>
    void flush_simple(std::vector<Op>& batch, std::vector<Op>& store, int kind, std::vector<Launch>& out, int level)
    {
        if (batch.empty())
            return;
        Launch L{};
        L.kind = kind;
        L.op_begin = (int64_t)store.size();
        L.op_count = (int32_t)batch.size();
        L.level = level;
        L.total_tiles = tiles_hint(batch);
        store.insert(store.end(), batch.begin(), batch.end());
        out.push_back(L);
        batch.clear();
    }

    // CTAs per op for the looping helper kernels = the largest per-op tile count of the batch
    static int tiles_hint(const std::vector<TransposeOp>& b)
    {
        int t = 1;
        for (auto& o : b)
            t = std::max(t, cdiv(o.rows, 32) * cdiv(o.cols, 32));
        return t;
    }
    template <class Op>
    static int tiles_hint(const std::vector<Op>&)
    {
        return 0;
    }

    // W = L11^-1 and Wt = W^T by block doubling, from the pivot-tile inverses the factorisation left on the
    // diagonal: pairs of finished bw x bw diagonal blocks (the second one may be shorter) are joined,
    // W21 = -W22 (L21 W11),  as  Tt = Wt11 L21^T  (parked in the unused upper part of W),  W21 = -W22 Tt^T,
    // Wt12 = W21^T.  Every pair of every front of the level is in the same three launches: 3 log2(k / 128)
    // launches per level.
    void build_trtri(const std::vector<uint32_t>& fl, std::vector<Launch>& out, int level)
    {
        std::vector<TransposeOp> trb;
        int maxk = 0;
        for (uint32_t fi : fl)
            maxk = std::max<int>(maxk, (int)s.fronts[fi].k);
        for (int bw = NB; bw < maxk; bw <<= 1) {
            std::vector<GemmOp> ga, gbb;
            for (uint32_t fi : fl) {
                const Front& f = s.fronts[fi];
                const uint32_t ldw = ldw_of(f);
                double* W = Wof(fi);
                double* Wt = Wtof(fi);
                for (int r0 = 0; r0 + bw < (int)f.k; r0 += 2 * bw) {
                    const int r1 = r0 + bw;
                    const int rows2 = std::min<int>(bw, (int)f.k - r1);
                    double* Tt = W + (size_t)r0 * ldw + r1;                    // bw x rows2
                    double* W21 = W + (size_t)r1 * ldw + r0;                   // rows2 x bw
                    share.clear();
                    add_gemm(ga, Wt + (size_t)r0 * ldw + r0, ldw, panel(f) + (size_t)r1 * f.ldk + r0, f.ldk, Tt, ldw, bw, rows2,
                             bw, GEMM_KLO_ROW);
                    share.clear();
                    add_gemm(gbb, W + (size_t)r1 * ldw + r1, ldw, Tt, ldw, W21, ldw, rows2, bw, rows2, GEMM_NEG | GEMM_KHI_ROW);
                    share.clear();
                    TransposeOp t{};
                    t.src = W21;
                    t.dst = Wt + (size_t)r0 * ldw + r1;
                    t.lds = ldw;
                    t.ldd = ldw;
                    t.rows = rows2;
                    t.cols = bw;
                    trb.push_back(t);
                }
            }
            flush_gemm(ga, out, level, T_TRTRI_A);
            flush_gemm(gbb, out, level, T_TRTRI_B);
            flush_simple(trb, p.transpose, L_TRANSPOSE, out, level);
        }
    }

    // ---- numeric factorisation ------------------------------------------------
    // schur: also the Schur updates of the fronts (a multi-GPU run does those of the top fronts apart, shared out)
    void factor_level(const std::vector<uint32_t>& fl, int lv, bool schur = true)
    {
        std::vector<GemmOp> gb;
        std::vector<DiagOp> db;
        int nsteps = 0;
        for (uint32_t f : fl)
            nsteps = std::max(nsteps, cdiv((int)s.fronts[f].k, NB));
        // Levels with enough fronts to fill the GPU from one block column per front run left-looking
        // (each block column is updated once, from all columns to its left, K = jb: half the panel traffic
        // and long K loops); sparse top levels run right-looking (2-D tile parallelism inside few big fronts).
        size_t row_tiles = 0;
        for (uint32_t f : fl)
            row_tiles += (size_t)cdiv((int)s.fronts[f].m, TILE_M);
        const bool left = row_tiles >= 2 * 148;
        for (int j = 0; j < nsteps; ++j) {
            const int jb = j * NB;
            if (left && jb > 0) {
                for (size_t i = 0; i < fl.size(); ++i) {
                    const Front& f = s.fronts[fl[i]];
                    if (jb >= (int)f.k)
                        continue;
                    int w = std::min<int>(NB, (int)f.k - jb);
                    double* A = panel(f) + (size_t)jb * f.ldk;
                    add_gemm(gb, A, f.ldk, A, f.ldk, A + jb, f.ldk, (int)f.m - jb, w, jb, GEMM_ACCUM | GEMM_NEG | GEMM_LOWER);
                }
                flush_gemm(gb, p.factor, lv, T_LEFT_UPDATE);
            }
            // pivot tiles
            for (size_t i = 0; i < fl.size(); ++i) {
                const Front& f = s.fronts[fl[i]];
                if (jb >= (int)f.k)
                    continue;
                DiagOp d{};
                d.D = panel(f) + (size_t)jb * f.ldk + jb;
                d.ldd = f.ldk;
                d.w = std::min<int>(NB, (int)f.k - jb);
                d.factor = 1;
                d.ldw = d.ldwt = ldw_of(f);
                d.W = Wof(fl[i]) + (size_t)jb * d.ldw + jb;
                d.Wt = Wtof(fl[i]) + (size_t)jb * d.ldw + jb;
                d.front = (int32_t)fl[i];
                db.push_back(d);
            }
            flush_diag(db, p.factor, lv);
            // rows below the pivot tile: P <- P * W^T
            for (size_t i = 0; i < fl.size(); ++i) {
                const Front& f = s.fronts[fl[i]];
                if (jb >= (int)f.k)
                    continue;
                int w = std::min<int>(NB, (int)f.k - jb);
                double* A = panel(f) + (size_t)(jb + w) * f.ldk + jb;
                add_gemm(gb, A, f.ldk, Wof(fl[i]) + (size_t)jb * ldw_of(f) + jb, ldw_of(f), A, f.ldk, (int)f.m - (jb + w), w, w,
                         0);
            }
            flush_gemm(gb, p.factor, lv, T_PANEL);
            if (left)
                continue;
            // trailing update inside the panel (columns still to be factorised)
            for (size_t i = 0; i < fl.size(); ++i) {
                const Front& f = s.fronts[fl[i]];
                if (jb >= (int)f.k)
                    continue;
                int w = std::min<int>(NB, (int)f.k - jb);
                int nc = (int)f.k - (jb + w);
                double* A = panel(f) + (size_t)(jb + w) * f.ldk + jb;
                double* C = panel(f) + (size_t)(jb + w) * f.ldk + (jb + w);
                add_gemm(gb, A, f.ldk, A, f.ldk, C, f.ldk, (int)f.m - (jb + w), nc, w, GEMM_ACCUM | GEMM_NEG | GEMM_LOWER);
            }
            flush_gemm(gb, p.factor, lv, T_RIGHT_UPDATE);
        }
        build_trtri(fl, p.factor, lv);
        if (schur)
            schur_level(fl, lv, false);
    }

    // Schur update -L21 L21^T of every front, one lower-triangular r x r product per front, scattered into the
    // ancestors' panels through the per-column target table (multi-GPU: into this rank's replicas of the top fronts —
    // the ranks' partial sums meet in the all-reduce before the ancestors' level).  dist: the tiles of the products are
    // shared out among the ranks (top fronts, whose factor every rank holds by then).
    void schur_level(const std::vector<uint32_t>& fl, int lv, bool dist)
    {
        std::vector<GemmOp> gb;
        for (uint32_t fi : fl) {
            const Front& f = s.fronts[fi];
            if (!f.r)
                continue;
            double* A = panel(f) + (size_t)f.k * f.ldk;
            add_gemm(gb, A, f.ldk, A, f.ldk, nullptr, 0, (int)f.r, (int)f.r, (int)f.k, GEMM_SCATTER | GEMM_NEG | GEMM_LOWER, 0,
                     b.coltgt + f.bnd_begin, Share{dist ? D_SUM : D_ALL, 0});
        }
        flush_gemm(gb, p.factor, lv, T_SCHUR);
    }

    void build_factor()
    {
        for (size_t lv = 0; lv < s.levels.size(); ++lv)
            factor_level(local_fronts(lv), (int)lv);
        if (multi())
            for (size_t lv = 0; lv < s.levels.size(); ++lv) {
                const std::vector<uint32_t> tl = top_fronts(lv);
                if (tl.empty())
                    continue;
                // the ranks' partial sums of these fronts (their subtrees' and the lower top fronts' Schur updates, the
                // assembled blocks on one rank) become the assembled fronts on every rank
                std::vector<ReduceOp> ranges;
                for (uint32_t fi : tl)
                    ranges.push_back(ReduceOp{s.fronts[fi].panel_off, (uint64_t)s.fronts[fi].m * s.fronts[fi].ldk});
                add_allreduce(p.factor, (int)lv, MC_PANELS, ranges);
                // Each top front is factorised by one rank (the fronts of a level by different ranks, side by side) with the
                // one-GPU launch lists, and the factor — panel, W, Wt — pushed into every replica: every rank needs it for the
                // substitutions and for its share of the selected inverse.  (Sharing the factorisation itself out tile by
                // tile was measured first: ~100 dependent pivot steps of two barriers each cost more than they saved,
                // profiles/r2_bench_8gpu_c4_first.json.)
                std::vector<uint32_t> mine;
                for (uint32_t fi : tl)
                    if (s.fronts[fi].owner == s.rank)
                        mine.push_back(fi);
                factor_level(mine, (int)lv, false);
                for (int stage = 0; stage < 2; ++stage) {
                    for (uint32_t fi : tl) {
                        const Front& f = s.fronts[fi];
                        add_broadcast(stage, f.owner, MC_PANELS, panel(f), f.ldk, (int)f.m, (int)f.k);
                        add_broadcast(stage, f.owner, MC_WBUF, Wof(fi), ldw_of(f), (int)f.k, (int)f.k);
                        add_broadcast(stage, f.owner, MC_WBUF, Wtof(fi), ldw_of(f), (int)f.k, (int)f.k);
                    }
                    flush_push(p.factor, (int)lv);
                }
                // the Schur updates of these fronts (the larger part of their factorisation flops) by all ranks together, each
                // scattering its tiles into its own replicas of the fronts above: partial sums again, joined before their level
                schur_level(tl, (int)lv, true);
            }
        for (auto& L : p.factor)
            p.factor_flops += L.flops;
    }

    // ---- triangular solves ----------------------------------------------------
    // x holds the right-hand side and, at the end, the solution; y the forward-substituted vector.
    //   forward  (bottom-up):  y1 = W x1;            x[boundary rows] -= L21 y1
    //   backward (top-down):   y1 -= L21^T x[boundary rows];   x1 = Wt y1
    void add_trimv(std::vector<TrimvOp>& tb, uint32_t fi, bool upper)
    {
        const Front& f = s.fronts[fi];
        const bool wide = (int)f.k >= TRIMV_WIDE_K;
        const int ROWS = wide ? TRIMV_WIDE_ROWS : 64;
        const uint32_t ldw = ldw_of(f);
        const double* A = upper ? Wtof(fi) : Wof(fi);
        for (int r0 = 0; r0 < (int)f.k; r0 += ROWS) {
            TrimvOp t{};
            t.A = A + (size_t)r0 * ldw;
            t.ld = ldw;
            t.row0 = r0;
            t.nrows = std::min<int>(ROWS, (int)f.k - r0);
            t.k = (int32_t)f.k;
            t.upper = upper ? 1 : 0;
            t.wide = wide ? 1 : 0;
            t.x = (upper ? b.y : b.x) + 3 * (size_t)f.own_begin;
            t.y = (upper ? b.x : b.y) + 3 * (size_t)f.own_begin;
            tb.push_back(t);
        }
    }
    void add_gemv(std::vector<GemvOp>& vb, uint32_t fi)
    {
        const int CHUNK = 256;
        const Front& f = s.fronts[fi];
        for (int jb = 0; jb < (int)f.k; jb += NB)
            for (int r0 = (int)f.k; r0 < (int)f.m; r0 += CHUNK) {
                GemvOp g{};
                g.P = panel(f) + (size_t)r0 * f.ldk + jb;
                g.rowidx = b.rowidx + p.rowidx_off[fi] + r0;
                g.xj = b.y + 3 * (size_t)f.own_begin + jb;
                g.ld = f.ldk;
                g.nrows = std::min<int>(CHUNK, (int)f.m - r0);
                g.w = std::min<int>(NB, (int)f.k - jb);
                vb.push_back(g);
            }
    }
    void build_solves()
    {
        std::vector<TrimvOp> tb;
        std::vector<GemvOp> vb;
        auto forward = [&](const std::vector<uint32_t>& fl, int lv) {
            for (uint32_t fi : fl) {
                add_trimv(tb, fi, false);
                add_gemv(vb, fi);
            }
            flush_simple(tb, p.tri, L_TRI_FWD, p.fwd, lv);
            flush_simple(vb, p.gemv, L_GEMV_FWD, p.fwd, lv);
        };
        auto backward = [&](const std::vector<uint32_t>& fl, int lv) {
            for (uint32_t fi : fl) {
                add_gemv(vb, fi);
                add_trimv(tb, fi, true);
            }
            flush_simple(vb, p.gemv, L_GEMV_BWD, p.bwd, lv);
            flush_simple(tb, p.tri, L_TRI_BWD, p.bwd, lv);
        };
        for (size_t lv = 0; lv < s.levels.size(); ++lv)
            forward(local_fronts(lv), (int)lv);
        if (multi()) {
            // the subtrees' contributions to the top fronts' right-hand sides are summed over the ranks; from there on every
            // rank carries the (cheap) substitutions of the replicated top fronts itself and needs no further exchange
            std::vector<ReduceOp> ranges;
            for (const Front& f : s.fronts)
                if (f.top)
                    ranges.push_back(ReduceOp{3ull * f.own_begin, f.k});
            add_allreduce(p.fwd, s.cut_level, MC_X, ranges);
            for (size_t lv = 0; lv < s.levels.size(); ++lv)
                forward(top_fronts(lv), (int)lv);
            for (size_t lvi = s.levels.size(); lvi-- > 0;)
                backward(top_fronts(lvi), (int)lvi);
        }
        for (size_t lvi = s.levels.size(); lvi-- > 0;)
            backward(local_fronts(lvi), (int)lvi);
    }

    // ---- selected inverse -------------------------------------------------------
    // dist (multi-GPU): the replicated top fronts of one level, their workspaces at the same pool offsets on every rank;
    // the tiles of the four products are shared out among the ranks, every rank pushes the tiles it has finished into
    // all replicas and the ranks meet before the next product reads them.
    void build_selinv_chunk(const std::vector<uint32_t>& chunk, const std::vector<size_t>& base, int level, bool dist)
    {
        std::vector<GemmOp> gb;
        std::vector<GatherOp> gab;
        const int D = dist ? D_SUM : D_ALL;
        // no clearing of the workspace: every tile that is read has been written before (the K-range
        // flags keep the triangular products inside the written tiles)
        for (size_t i = 0; i < chunk.size(); ++i) {
            const Front& f = s.fronts[chunk[i]];
            SelinvWs w = ws_layout(f);
            double* ws = b.pool + base[i];
            for (uint32_t ti = 0; ti < f.tgt_count; ++ti) {
                const Target& tg = s.targets[f.tgt_begin + ti];
                const Front& an = s.fronts[tg.anc];
                GatherOp g{};
                g.Z = panel(an);
                g.ld = an.ldk;
                g.rowmap = b.rowmap + tg.rowmap_off;
                g.G = ws + w.G;
                g.ldg = w.ldg;
                g.jb = (int32_t)tg.jb;
                g.je = (int32_t)tg.je;
                g.nb = (int32_t)f.bnd_count;
                g.col0 = (int32_t)tg.col0;
                gab.push_back(g);
            }
        }
        if (!gab.empty()) {
            // the tiles (16 x 16 stations) on or below the diagonal of every op, as one work list for the launch
            Launch L{};
            L.kind = L_GATHER;
            L.level = level;
            L.op_begin = (int64_t)p.gather.size();
            L.op_count = (int32_t)gab.size();
            L.tile_begin = (int64_t)p.gather_tiles.size();
            append_gather_tiles(gab, p.gather_tiles);
            L.total_tiles = (int32_t)(p.gather_tiles.size() - (size_t)L.tile_begin);
            p.gather.insert(p.gather.end(), gab.begin(), gab.end());
            p.selinv.push_back(L);
            gab.clear();
        }
        // Yt = Wt * L21^T
        for (size_t i = 0; i < chunk.size(); ++i) {
            const Front& f = s.fronts[chunk[i]];
            if (!f.r)
                continue;
            SelinvWs w = ws_layout(f);
            double* ws = b.pool + base[i];
            add_gemm(gb, Wtof(chunk[i]), ldw_of(f), panel(f) + (size_t)f.k * f.ldk, f.ldk, ws + w.Yt, w.ldr, (int)f.k, (int)f.r,
                     (int)f.k, GEMM_KLO_ROW, 0, nullptr, Share{D, dist ? MC_POOL : 0});
        }
        flush_gemm(gb, p.selinv, level, T_YT);
        if (dist)
            flush_push(p.selinv, level);
        // Z21 = -G * Y   (overwrites L21), with its transpose Z21t stored by the same epilogue
        for (size_t i = 0; i < chunk.size(); ++i) {
            const Front& f = s.fronts[chunk[i]];
            if (!f.r)
                continue;
            SelinvWs w = ws_layout(f);
            double* ws = b.pool + base[i];
            add_gemm(gb, ws + w.G, w.ldg, ws + w.Yt, w.ldr, panel(f) + (size_t)f.k * f.ldk, f.ldk, (int)f.r, (int)f.k,
                     (int)f.r, GEMM_NEG | GEMM_DUAL, 0, nullptr, Share{D, dist ? (MC_PANELS | (MC_POOL << 8)) : 0});
            gb.back().Ct = ws + w.Z21t;
            gb.back().ldct = w.ldr;
        }
        flush_gemm(gb, p.selinv, level, T_Z21);
        if (dist)
            flush_push(p.selinv, level);
        // Z11 = Wt Wt^T   (overwrites L11, lower triangle); a front without a boundary is finished by this product
        for (size_t i = 0; i < chunk.size(); ++i) {
            const Front& f = s.fronts[chunk[i]];
            add_gemm(gb, Wtof(chunk[i]), ldw_of(f), Wtof(chunk[i]), ldw_of(f), panel(f), f.ldk, (int)f.k, (int)f.k, (int)f.k,
                     GEMM_LOWER | GEMM_KLO_MAX, 0, nullptr, Share{D, (dist && f.r == 0) ? MC_PANELS : 0});
        }
        flush_gemm(gb, p.selinv, level, T_Z11_WW);
        // Z11 -= Yt * Z21t^T
        for (size_t i = 0; i < chunk.size(); ++i) {
            const Front& f = s.fronts[chunk[i]];
            if (!f.r)
                continue;
            SelinvWs w = ws_layout(f);
            double* ws = b.pool + base[i];
            add_gemm(gb, ws + w.Yt, w.ldr, ws + w.Z21t, w.ldr, panel(f), f.ldk, (int)f.k, (int)f.k, (int)f.r,
                     GEMM_ACCUM | GEMM_NEG | GEMM_LOWER, 0, nullptr, Share{D, dist ? MC_PANELS : 0});
        }
        flush_gemm(gb, p.selinv, level, T_Z11_YZ);
        if (dist)
            flush_push(p.selinv, level);
    }

    void build_selinv()
    {
        if (multi())
            for (size_t lvi = s.levels.size(); lvi-- > 0;) {
                // all top fronts of the level together (min_pool_doubles has made room for them on every rank)
                const std::vector<uint32_t> tl = top_fronts(lvi);
                if (tl.empty())
                    continue;
                std::vector<size_t> base;
                size_t used = 0;
                for (uint32_t fi : tl) {
                    base.push_back(used);
                    used += ws_layout(s.fronts[fi]).total;
                }
                if (used > b.pool_doubles) {
                    err = "workspace pool smaller than the top fronts of a level";
                    return;
                }
                build_selinv_chunk(tl, base, (int)lvi, true);
            }
        for (size_t lvi = s.levels.size(); lvi-- > 0;) {
            const std::vector<uint32_t> fl = local_fronts(lvi);
            std::vector<uint32_t> chunk;
            std::vector<size_t> base;
            size_t used = 0;
            for (uint32_t fi : fl) {
                size_t need = ws_layout(s.fronts[fi]).total;
                if (need > b.pool_doubles) {
                    err = "workspace pool smaller than the largest front";
                    return;
                }
                if (used + need > b.pool_doubles) {
                    build_selinv_chunk(chunk, base, (int)lvi, false);
                    chunk.clear();
                    base.clear();
                    used = 0;
                }
                chunk.push_back(fi);
                base.push_back(used);
                used += need;
            }
            if (!chunk.empty())
                build_selinv_chunk(chunk, base, (int)lvi, false);
        }
        for (auto& L : p.selinv)
            p.selinv_flops += L.flops;
    }
};

}  // namespace

size_t selinv_workspace(const Front& f) { return ws_layout(f).total; }

size_t min_pool_doubles(const Symbolic& s)
{
    size_t need = 0;
    for (const Front& f : s.fronts)
        if (f.owner == s.rank && !f.top)
            need = std::max(need, ws_layout(f).total);
    // the replicated top fronts of a level are inverted together, their workspaces side by side
    for (auto& lv : s.levels) {
        size_t t = 0;
        for (uint32_t f : lv)
            if (s.fronts[f].top)
                t += ws_layout(s.fronts[f]).total;
        need = std::max(need, t);
    }
    return std::max<size_t>(16, need);
}

size_t wbuf_doubles(const Symbolic& s)
{
    size_t o = 0;
    for (const Front& f : s.fronts)
        if (f.owner == s.rank || f.top)
            o += 2 * wblock(f);
    return std::max<size_t>(16, o);
}

size_t ideal_pool_doubles(const Symbolic& s)
{
    size_t best = min_pool_doubles(s);
    for (auto& lv : s.levels) {
        size_t t = 0;
        for (uint32_t f : lv)
            if (s.fronts[f].owner == s.rank && !s.fronts[f].top)
                t += ws_layout(s.fronts[f]).total;
        best = std::max(best, t);
    }
    return best;
}

void build_rowidx(const Symbolic& s, Plan& p)
{
    p.rowidx_off.assign(s.fronts.size(), 0);
    size_t tot = 0;
    for (size_t f = 0; f < s.fronts.size(); ++f) {
        p.rowidx_off[f] = tot;
        tot += s.fronts[f].m;
    }
    p.rowidx.resize(tot);
    for (size_t fi = 0; fi < s.fronts.size(); ++fi) {
        const Front& f = s.fronts[fi];
        int32_t* r = p.rowidx.data() + p.rowidx_off[fi];
        for (uint32_t i = 0; i < f.k; ++i)
            r[i] = (int32_t)(3 * f.own_begin + i);
        for (uint32_t i = 0; i < f.bnd_count; ++i)
            for (int c = 0; c < 3; ++c)
                r[f.k + 3 * i + c] = (int32_t)(3 * s.bnd[f.bnd_begin + i] + c);
    }
}

std::string build_plan(const Symbolic& s, const PlanBuffers& b, Plan& p)
{
    p.gemm.clear();
    p.tiles.clear();
    p.diag.clear();
    p.tri.clear();
    p.gemv.clear();
    p.transpose.clear();
    p.gather.clear();
    p.gather_tiles.clear();
    p.reduce.clear();
    p.push.clear();
    p.factor.clear();
    p.fwd.clear();
    p.bwd.clear();
    p.selinv.clear();
    p.factor_flops = p.selinv_flops = 0;
    p.nvlink_read_bytes = p.nvlink_write_bytes = 0;
    p.barriers = 0;
    p.launches_tile64 = 0;
    if (b.pool_doubles < min_pool_doubles(s))
        return "workspace pool smaller than the largest front";
    // scatter tables of the Schur updates
    p.tgt.assign(s.targets.size(), ScatterTarget{});
    p.coltgt.assign(s.bnd.size(), -1);
    for (const Front& f : s.fronts) {
        if (f.owner != s.rank && !f.top)
            continue;
        for (uint32_t t = 0; t < f.tgt_count; ++t) {
            const Target& tg = s.targets[f.tgt_begin + t];
            const Front& an = s.fronts[tg.anc];
            ScatterTarget& st = p.tgt[f.tgt_begin + t];
            st.C = b.panels + an.panel_off;
            st.rowmap = b.rowmap + tg.rowmap_off;
            st.ldc = an.ldk;
            st.jb = (int32_t)tg.jb;
            for (uint32_t i = tg.jb; i < tg.je; ++i)
                p.coltgt[f.bnd_begin + i] = (int32_t)(f.tgt_begin + t);
        }
    }
    Builder B(s, b, p);
    B.build_factor();
    B.build_solves();
    B.build_selinv();
    return B.err;
}

}  // namespace gadj
