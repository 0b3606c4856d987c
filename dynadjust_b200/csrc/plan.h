// plan.h — turns the symbolic structure into static launch lists (host).
#pragma once
#include <string>
#include <vector>

#include "kernels.h"
#include "symbolic.h"

namespace gadj {

enum LaunchKind : int32_t {
    L_GEMM = 0,
    L_DIAG,
    L_TRI_FWD,
    L_TRI_BWD,
    L_GEMV_FWD,
    L_GEMV_BWD,
    L_TRANSPOSE,
    L_GATHER,
    L_ZERO,
    L_BARRIER,   // multi-GPU: all ranks meet (device-side barrier over NVLink); stores into peers' replicas are visible after it
    L_ALLREDUCE, // multi-GPU: sum of the ranks' replicas of some ranges of a buffer, written back to every replica
    L_PUSH,      // multi-GPU: blocks this rank has finished, copied into every peer's replica
};

struct Launch {
    int32_t kind;
    int32_t op_count;
    int64_t op_begin;     // index into the per-kind op array
    int32_t total_tiles;  // L_GEMM / L_GATHER: tiles with work (entries of Plan::tiles / gather_tiles from tile_begin); transpose: CTAs per op
    int64_t tile_begin;   // L_GEMM, L_GATHER
    int32_t level;
    double* zero_ptr;     // L_ZERO
    size_t zero_bytes;
    double flops;         // algorithmic flops of this launch (GEMM: 2MNK, halved for LOWER)
    int32_t tag;          // which step of the algorithm (profiling label)
    int32_t buf;          // L_ALLREDUCE: McBuf of the buffer the ranges live in
    int32_t shape;        // L_GEMM: TileShape of the launch's tiles
};

enum LaunchTag : int32_t {
    T_NONE = 0, T_LEFT_UPDATE, T_PANEL, T_RIGHT_UPDATE, T_SCHUR, T_TRTRI_A, T_TRTRI_B, T_YT, T_Z21, T_Z11_WW, T_Z11_YZ,
};

struct PlanBuffers {
    double* panels = nullptr;     // device, Symbolic::panel_doubles
    double* pool = nullptr;       // device workspace pool
    size_t pool_doubles = 0;
    double* x = nullptr;          // device, 3*nstn, elimination order: right-hand side, then the solution
    double* y = nullptr;          // device, 3*nstn: the forward-substituted vector (second vector of the solves)
    double* wbuf = nullptr;       // device, wbuf_doubles(): W = L11^-1 and Wt = W^T of every front this rank owns
    const int32_t* rowmap = nullptr;  // device copy of Symbolic::rowmap
    const int32_t* rowidx = nullptr;  // device: global unknown index per front row (concatenated, Plan::rowidx_off)
    const ScatterTarget* tgt = nullptr;   // device, Symbolic::targets.size() entries (filled from Plan::tgt)
    const int32_t* coltgt = nullptr;      // device, Symbolic::bnd.size() entries (filled from Plan::coltgt)
    // GEMM tile shape: 0 = per launch by the padded-work model (below), 64 / 128 = that shape wherever it is allowed
    int32_t gemm_tile = 0;
};

struct Plan {
    std::vector<GemmOp> gemm;
    std::vector<GemmTile> tiles;
    std::vector<ScatterTarget> tgt;         // per Symbolic::targets entry (device pointers inside)
    std::vector<int32_t> coltgt;            // per Symbolic::bnd entry: index into tgt of the ancestor owning that station
    std::vector<DiagOp> diag;
    std::vector<TrimvOp> tri;
    std::vector<GemvOp> gemv;
    std::vector<TransposeOp> transpose;
    std::vector<GatherOp> gather;
    std::vector<GatherTile> gather_tiles;   // work lists of the gather launches (Launch::tile_begin / total_tiles)
    std::vector<ReduceOp> reduce;
    std::vector<PushOp> push;
    std::vector<Launch> factor, fwd, bwd, selinv;
    std::vector<int32_t> rowidx;            // host copy (uploaded by the caller before build_plan's pointers are used)
    std::vector<uint64_t> rowidx_off;       // per front
    double factor_flops = 0, selinv_flops = 0;
    // multi-GPU, per iteration of this rank: bytes read from / stored into the peers' replicas, barriers
    double nvlink_read_bytes = 0, nvlink_write_bytes = 0;
    uint64_t barriers = 0;
    uint64_t launches_tile64 = 0;   // GEMM launches planned with 64 x 64 tiles
    // device copies of the op arrays (owned by the context)
    GemmOp* d_gemm = nullptr;
    DiagOp* d_diag = nullptr;
    TrimvOp* d_tri = nullptr;
    GemvOp* d_gemv = nullptr;
    TransposeOp* d_transpose = nullptr;
    GatherOp* d_gather = nullptr;
};

// doubles needed for the persistent inverse pivot blocks (W and Wt of every owned front)
size_t wbuf_doubles(const Symbolic& s);
// per-front selected-inverse workspace need (doubles)
size_t selinv_workspace(const Front& f);
// smallest usable pool (doubles): the largest single front
size_t min_pool_doubles(const Symbolic& s);
// pool size that lets every level run as a single chunk
size_t ideal_pool_doubles(const Symbolic& s);

// the gather work list of a batch of ops: the 16 x 16-station tiles on or below each op's diagonal, appended to `out`
inline void append_gather_tiles(const std::vector<GatherOp>& ops, std::vector<GatherTile>& out)
{
    const int GT = GATHER_TILE_STATIONS;
    for (size_t o = 0; o < ops.size(); ++o) {
        const int ni = ops[o].nb - ops[o].jb, nj = ops[o].je - ops[o].jb;
        for (int ti = 0; ti < (ni + GT - 1) / GT; ++ti)
            for (int tj = 0; tj < (nj + GT - 1) / GT && tj * GT <= ti * GT + GT - 1; ++tj)
                out.push_back(GatherTile{(int32_t)o, (uint16_t)ti, (uint16_t)tj});
    }
}

void build_rowidx(const Symbolic& s, Plan& p);
// builds all four launch lists; encodes TMA descriptors through dev::encode_tma_2d
std::string build_plan(const Symbolic& s, const PlanBuffers& b, Plan& p);

}  // namespace gadj
