// symbolic.h — host-side symbolic analysis for the supernodal block Cholesky.
//
// Replaces, structurally, the reference's block bookkeeping for phased adjustment
// (v_ISL_/v_JSL_/v_blockStationsMap_, ADJH:1112-1123, ADJH:1216-1218, seg_file.cpp:432-486):
// a "front" is one supernode = the reference's block, its own stations are the
// block's inner stations (ISL) and its boundary stations are the junction
// stations (JSL).  The tree is either the .seg chain handed in by the caller or
// a geometric nested dissection computed here.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace gadj {

struct Front {
    uint32_t own_begin = 0;   // first elimination position owned
    uint32_t own_count = 0;   // stations owned (inner stations)
    uint32_t bnd_begin = 0;   // offset into Symbolic::bnd (boundary station positions, ascending)
    uint32_t bnd_count = 0;
    int32_t parent = -1;
    int32_t level = 0;        // height above the leaves
    uint32_t k = 0, r = 0, m = 0;  // unknowns: own, boundary, total
    uint32_t ldk = 0;         // panel row pitch (doubles), even
    uint64_t panel_off = 0;   // offset (doubles) of the m x ldk row-major panel
    uint32_t tgt_begin = 0, tgt_count = 0;  // range into Symbolic::targets
    int32_t owner = 0;        // rank that holds this front (multi-GPU sharding by subtree); top fronts: the rank that assembles N into it
    uint8_t top = 0;          // 1: above the subtree cut — replicated on every rank, its tiles shared out among the ranks
    double work = 0;          // factor + inverse flops of this front
};

constexpr uint64_t NO_DEST = ~0ull;

// Update target: boundary stations [jb, je) of front `src` are owned by ancestor `anc`.
struct Target {
    uint32_t anc;        // ancestor front
    uint32_t jb, je;     // range in src's boundary list (station level)
    uint32_t col0;       // first own column (station level) in anc of boundary station jb
    uint64_t rowmap_off; // offset into Symbolic::rowmap; entry i-jb = local station row in anc of boundary station i (i in [jb, bnd_count))
};

struct Symbolic {
    uint32_t nstn = 0;
    std::vector<uint32_t> pos_of_stn;   // station index -> elimination position
    std::vector<uint32_t> stn_of_pos;   // inverse
    std::vector<uint32_t> front_of_pos; // position -> front
    std::vector<Front> fronts;          // in elimination order (children before parents)
    std::vector<uint32_t> bnd;          // concatenated boundary lists (positions)
    std::vector<Target> targets;
    std::vector<int32_t> rowmap;        // concatenated station-level row maps
    std::vector<std::vector<uint32_t>> levels;  // fronts by level
    uint64_t panel_doubles = 0;
    uint64_t top_panel_doubles = 0;     // leading part of the panel storage: the replicated top fronts (same layout on every rank)
    // block-CSR pattern of N in elimination order: column block p holds the diagonal slot
    // followed by its later neighbours q > p (ascending)
    std::vector<uint64_t> ncol_ptr;     // size nstn+1, slot offsets
    std::vector<uint32_t> nrow;         // row block position per slot
    std::vector<uint64_t> ndest;        // destination offset (doubles) of the slot's (0,0) element in panel storage
    std::vector<uint32_t> ndest_ld;     // row pitch at the destination
    double factor_flops = 0, inverse_flops = 0;
    uint64_t nnz_l_blocks = 0;
    // sharding (world == 1: every front owned by rank 0, nothing is "top")
    int32_t world = 1, rank = 0;
    int32_t cut_level = 1 << 30;               // first level that holds a top front
    std::vector<uint8_t> pos_owned;             // per elimination position: its front is owned by this rank
    double my_factor_flops = 0, my_inverse_flops = 0;
    std::vector<double> rank_load;              // per rank: work (flops) of the fronts in its subtrees
};

struct OrderingOptions {
    uint32_t leaf_stations = 96;      // nested-dissection leaf size
    uint32_t cut_degree_to_sep = 8;   // vertices with this many cut edges go straight to the separator
    bool dense = false;               // single front holding every station (BASELINE config C2 "dense normals")
};

// edges: undirected station pairs (duplicates / self loops tolerated).
// lat/lon (radians) drive the geometric dissection.
// blocks (optional): chain segmentation as in a .seg file — block b's inner stations isl[isl_off[b]..isl_off[b+1]).
// Returns empty string on success, else an error message.
std::string analyse(uint32_t nstn, const std::vector<std::pair<uint32_t, uint32_t>>& edges,
                    const double* lat, const double* lon, const OrderingOptions& opt,
                    uint32_t nblocks, const uint32_t* isl_off, const uint32_t* isl,
                    Symbolic& out);

// Assign fronts to ranks by subtree (largest subtrees split until there are >= world of them, then
// longest-processing-time packing), lay out this rank's panels (owned fronts + every top front) and compute
// the normal-matrix destinations (NO_DEST for blocks whose front another rank assembles).
void finalize_layout(Symbolic& s, int world, int rank);

// slot of block (row position q, column position p), q >= p; UINT64_MAX when absent
uint64_t find_slot(const Symbolic& s, uint32_t q, uint32_t p);

}  // namespace gadj
