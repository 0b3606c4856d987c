// symbolic.cpp — ordering, supernode tree, front layout and scatter maps (host).
#include "symbolic.h"

#include "par.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdlib>
#include <numeric>
#include <thread>

namespace gadj {
namespace {

struct Graph {
    std::vector<uint64_t> ptr;
    std::vector<uint32_t> adj;
};

Graph build_graph(uint32_t n, const std::vector<std::pair<uint32_t, uint32_t>>& edges)
{
    Graph g;
    g.ptr.assign((size_t)n + 1, 0);
    for (auto& e : edges) {
        if (e.first == e.second)
            continue;
        g.ptr[e.first + 1]++;
        g.ptr[e.second + 1]++;
    }
    for (uint32_t i = 0; i < n; ++i)
        g.ptr[i + 1] += g.ptr[i];
    std::vector<uint32_t> tmp(g.ptr[n]);
    std::vector<uint64_t> fill(g.ptr.begin(), g.ptr.end() - 1);
    for (auto& e : edges) {
        if (e.first == e.second)
            continue;
        tmp[fill[e.first]++] = e.second;
        tmp[fill[e.second]++] = e.first;
    }
    // sort + unique per vertex (all host threads), then compact
    std::vector<uint64_t> nptr((size_t)n + 1, 0);
    parallel_for(n, [&](uint64_t v0, uint64_t v1) {
        for (uint64_t i = v0; i < v1; ++i) {
            auto b = tmp.begin() + g.ptr[i], e = tmp.begin() + g.ptr[i + 1];
            std::sort(b, e);
            nptr[i + 1] = (uint64_t)(std::unique(b, e) - b);
        }
    });
    for (uint32_t i = 0; i < n; ++i)
        nptr[i + 1] += nptr[i];
    g.adj.resize(nptr[n]);
    parallel_for(n, [&](uint64_t v0, uint64_t v1) {
        for (uint64_t i = v0; i < v1; ++i)
            std::copy(tmp.begin() + g.ptr[i], tmp.begin() + g.ptr[i] + (nptr[i + 1] - nptr[i]), g.adj.begin() + nptr[i]);
    });
    g.ptr.swap(nptr);
    return g;
}

// ---- geometric nested dissection ------------------------------------------
struct Dissector {
    const Graph& g;
    const double* lat;
    const double* lon;
    const OrderingOptions& opt;
    std::vector<uint32_t> label;
    std::atomic<uint32_t> next_tag{1};
    typedef std::vector<std::vector<uint32_t>> Nodes;   // supernodes in elimination order

    Dissector(const Graph& G, const double* la, const double* lo, const OrderingOptions& o)
        : g(G), lat(la), lon(lo), opt(o), label(G.ptr.size() - 1, 0)
    {
    }

    static void emit(Nodes& out, std::vector<uint32_t>& v)
    {
        if (v.empty())
            return;
        std::sort(v.begin(), v.end());
        out.emplace_back(std::move(v));
    }

    // the two halves of a split are independent: the first levels of the recursion run them on separate threads
    // (labels are written per vertex, tags come from an atomic counter)
    Nodes run(std::vector<uint32_t>& verts, int depth = 0)
    {
        Nodes out;
        if (verts.size() <= opt.leaf_stations) {
            emit(out, verts);
            return out;
        }
        // pick the wider axis (metres, roughly): x = lon * cos(mean lat), y = lat
        double la0 = 1e300, la1 = -1e300, lo0 = 1e300, lo1 = -1e300, lam = 0;
        for (uint32_t v : verts) {
            la0 = std::min(la0, lat[v]);
            la1 = std::max(la1, lat[v]);
            lo0 = std::min(lo0, lon[v]);
            lo1 = std::max(lo1, lon[v]);
            lam += lat[v];
        }
        lam /= (double)verts.size();
        bool split_lon = (lo1 - lo0) * std::cos(lam) > (la1 - la0);
        const double* key = split_lon ? lon : lat;
        size_t half = verts.size() / 2;
        std::nth_element(verts.begin(), verts.begin() + half, verts.end(), [&](uint32_t a, uint32_t b) {
            return key[a] < key[b] || (key[a] == key[b] && a < b);
        });
        const uint32_t tagL = next_tag.fetch_add(3), tagR = tagL + 1, tagS = tagL + 2;
        for (size_t i = 0; i < verts.size(); ++i)
            label[verts[i]] = i < half ? tagL : tagR;
        // vertices with many cut edges (hubs) go straight into the separator
        std::vector<uint32_t> S;
        for (uint32_t v : verts) {
            uint32_t other = label[v] == tagL ? tagR : tagL;
            uint32_t cd = 0;
            for (uint64_t e = g.ptr[v]; e < g.ptr[v + 1]; ++e)
                cd += label[g.adj[e]] == other;
            if (cd >= opt.cut_degree_to_sep)
                S.push_back(v);
        }
        for (uint32_t v : S)
            label[v] = tagS;
        // one-sided boundary of the smaller side closes the remaining cut edges
        std::vector<uint32_t> BL, BR;
        for (uint32_t v : verts) {
            if (label[v] == tagS)
                continue;
            uint32_t other = label[v] == tagL ? tagR : tagL;
            for (uint64_t e = g.ptr[v]; e < g.ptr[v + 1]; ++e)
                if (label[g.adj[e]] == other) {
                    (label[v] == tagL ? BL : BR).push_back(v);
                    break;
                }
        }
        std::vector<uint32_t>& B = BL.size() <= BR.size() ? BL : BR;
        for (uint32_t v : B) {
            label[v] = tagS;
            S.push_back(v);
        }
        std::vector<uint32_t> L, R;
        L.reserve(half);
        R.reserve(verts.size() - half);
        for (uint32_t v : verts) {
            if (label[v] == tagL)
                L.push_back(v);
            else if (label[v] == tagR)
                R.push_back(v);
        }
        if (L.empty() || R.empty() || S.size() * 2 > verts.size()) {
            // degenerate split (clique-like subgraph): keep it as one dense front
            emit(out, verts);
            return out;
        }
        std::vector<uint32_t>().swap(verts);
        Nodes right;
        if (depth < 4 && L.size() + R.size() > 20000) {
            std::thread other([&] { right = run(R, depth + 1); });
            out = run(L, depth + 1);
            other.join();
        } else {
            out = run(L, depth + 1);
            right = run(R, depth + 1);
        }
        for (auto& v : right)
            out.emplace_back(std::move(v));
        emit(out, S);
        return out;
    }
};

}  // namespace

uint64_t find_slot(const Symbolic& s, uint32_t q, uint32_t p)
{
    uint64_t b = s.ncol_ptr[p], e = s.ncol_ptr[p + 1];
    if (q == p)
        return b;
    auto it = std::lower_bound(s.nrow.begin() + b + 1, s.nrow.begin() + e, q);
    if (it == s.nrow.begin() + e || *it != q)
        return UINT64_MAX;
    return (uint64_t)(it - s.nrow.begin());
}

std::string analyse(uint32_t nstn, const std::vector<std::pair<uint32_t, uint32_t>>& edges, const double* lat,
                    const double* lon, const OrderingOptions& opt, uint32_t nblocks, const uint32_t* isl_off,
                    const uint32_t* isl, Symbolic& out)
{
    out = Symbolic();
    out.nstn = nstn;
    if (nstn == 0)
        return "no stations";
    for (auto& e : edges)
        if (e.first >= nstn || e.second >= nstn)
            return "measurement refers to a station index beyond the station list";
    Graph g = build_graph(nstn, edges);

    // ---- 1. partition into supernodes, in elimination order -----------------
    std::vector<std::vector<uint32_t>> sn;
    if (opt.dense) {
        sn.emplace_back(nstn);
        std::iota(sn[0].begin(), sn[0].end(), 0u);
    } else if (nblocks > 1) {
        std::vector<uint8_t> seen(nstn, 0);
        for (uint32_t b = 0; b < nblocks; ++b) {
            std::vector<uint32_t> v(isl + isl_off[b], isl + isl_off[b + 1]);
            for (uint32_t s : v) {
                if (s >= nstn)
                    return "block list refers to a station index beyond the station list";
                if (seen[s])
                    return "station appears as an inner station of two blocks";
                seen[s] = 1;
            }
            if (!v.empty())
                sn.emplace_back(std::move(v));
        }
        for (uint32_t s = 0; s < nstn; ++s)
            if (!seen[s])
                return "station is not an inner station of any block";
    } else {
        Dissector d(g, lat, lon, opt);
        std::vector<uint32_t> all(nstn);
        std::iota(all.begin(), all.end(), 0u);
        sn = d.run(all);
    }

    const uint32_t F = (uint32_t)sn.size();
    out.pos_of_stn.assign(nstn, 0);
    out.stn_of_pos.assign(nstn, 0);
    out.front_of_pos.assign(nstn, 0);
    out.fronts.assign(F, Front());
    {
        uint32_t p = 0;
        for (uint32_t f = 0; f < F; ++f) {
            out.fronts[f].own_begin = p;
            out.fronts[f].own_count = (uint32_t)sn[f].size();
            for (uint32_t s : sn[f]) {
                out.pos_of_stn[s] = p;
                out.stn_of_pos[p] = s;
                out.front_of_pos[p] = f;
                ++p;
            }
        }
        if (p != nstn)
            return "internal: ordering does not cover every station";
    }
    std::vector<std::vector<uint32_t>>().swap(sn);

    // ---- 2. supernodal symbolic factorisation --------------------------------
    std::vector<uint32_t> mark(nstn, UINT32_MAX);
    std::vector<std::vector<uint32_t>> children(F);
    std::vector<uint32_t> st;
    for (uint32_t f = 0; f < F; ++f) {
        Front& fr = out.fronts[f];
        const uint32_t own_end = fr.own_begin + fr.own_count;
        st.clear();
        for (uint32_t p = fr.own_begin; p < own_end; ++p) {
            uint32_t v = out.stn_of_pos[p];
            for (uint64_t e = g.ptr[v]; e < g.ptr[v + 1]; ++e) {
                uint32_t q = out.pos_of_stn[g.adj[e]];
                if (q >= own_end && mark[q] != f) {
                    mark[q] = f;
                    st.push_back(q);
                }
            }
        }
        for (uint32_t c : children[f]) {
            const Front& ch = out.fronts[c];
            for (uint32_t i = 0; i < ch.bnd_count; ++i) {
                uint32_t q = out.bnd[ch.bnd_begin + i];
                if (q >= own_end && mark[q] != f) {
                    mark[q] = f;
                    st.push_back(q);
                }
            }
        }
        std::sort(st.begin(), st.end());
        fr.bnd_begin = (uint32_t)out.bnd.size();
        fr.bnd_count = (uint32_t)st.size();
        out.bnd.insert(out.bnd.end(), st.begin(), st.end());
        if (!st.empty()) {
            fr.parent = (int32_t)out.front_of_pos[st[0]];
            children[fr.parent].push_back(f);
        }
    }

    // ---- 3. levels, dimensions, storage --------------------------------------
    int32_t maxlevel = 0;
    for (uint32_t f = 0; f < F; ++f) {
        Front& fr = out.fronts[f];
        int32_t lv = 0;
        for (uint32_t c : children[f])
            lv = std::max(lv, out.fronts[c].level + 1);
        fr.level = lv;
        maxlevel = std::max(maxlevel, lv);
        fr.k = 3 * fr.own_count;
        fr.r = 3 * fr.bnd_count;
        fr.m = fr.k + fr.r;
        fr.ldk = fr.k + (fr.k & 1u);
    }
    out.levels.assign((size_t)maxlevel + 1, {});
    for (uint32_t f = 0; f < F; ++f) {
        Front& fr = out.fronts[f];
        out.levels[fr.level].push_back(f);
        double k = fr.k, r = fr.r;
        double ff = k * k * k / 3.0 + k * k * r + k * r * r;
        double fi = 2.0 * k * k * k / 3.0 + 2.0 * k * k * r + 2.0 * k * r * r + 2.0 * k * k * r;
        fr.work = ff + fi;
        out.factor_flops += ff;
        out.inverse_flops += fi;
        out.nnz_l_blocks += (uint64_t)fr.own_count * (fr.own_count + 1) / 2 + (uint64_t)fr.own_count * fr.bnd_count;
    }

    // ---- 4. update targets and row maps --------------------------------------
    for (uint32_t f = 0; f < F; ++f) {
        Front& fr = out.fronts[f];
        fr.tgt_begin = (uint32_t)out.targets.size();
        const uint32_t* b = out.bnd.data() + fr.bnd_begin;
        uint32_t j = 0;
        while (j < fr.bnd_count) {
            uint32_t a = out.front_of_pos[b[j]];
            const Front& an = out.fronts[a];
            uint32_t a_end = an.own_begin + an.own_count;
            uint32_t je = j;
            while (je < fr.bnd_count && b[je] < a_end)
                ++je;
            Target t;
            t.anc = a;
            t.jb = j;
            t.je = je;
            t.col0 = b[j] - an.own_begin;
            t.rowmap_off = out.rowmap.size();
            const uint32_t* ab = out.bnd.data() + an.bnd_begin;
            uint32_t w = 0;
            for (uint32_t i = j; i < fr.bnd_count; ++i) {
                uint32_t p = b[i];
                if (p < a_end) {
                    out.rowmap.push_back((int32_t)(p - an.own_begin));
                } else {
                    while (w < an.bnd_count && ab[w] < p)
                        ++w;
                    if (w >= an.bnd_count || ab[w] != p)
                        return "internal: boundary station missing from an ancestor front";
                    out.rowmap.push_back((int32_t)(an.own_count + w));
                }
            }
            out.targets.push_back(t);
            j = je;
        }
        fr.tgt_count = (uint32_t)out.targets.size() - fr.tgt_begin;
    }

    // ---- 5. block pattern of N in elimination order + panel destinations -----
    out.ncol_ptr.assign((size_t)nstn + 1, 0);
    parallel_for(nstn, [&](uint64_t p0, uint64_t p1) {
        for (uint64_t p = p0; p < p1; ++p) {
            uint32_t v = out.stn_of_pos[p];
            uint64_t later = 0;
            for (uint64_t e = g.ptr[v]; e < g.ptr[v + 1]; ++e)
                later += out.pos_of_stn[g.adj[e]] > p;
            out.ncol_ptr[p + 1] = 1 + later;
        }
    });
    for (uint32_t p = 0; p < nstn; ++p)
        out.ncol_ptr[p + 1] += out.ncol_ptr[p];
    out.nrow.resize(out.ncol_ptr[nstn]);
    parallel_for(nstn, [&](uint64_t p0, uint64_t p1) {
        for (uint64_t p = p0; p < p1; ++p) {
            uint32_t v = out.stn_of_pos[p];
            uint64_t s = out.ncol_ptr[p];
            out.nrow[s++] = (uint32_t)p;
            for (uint64_t e = g.ptr[v]; e < g.ptr[v + 1]; ++e) {
                uint32_t q = out.pos_of_stn[g.adj[e]];
                if (q > p)
                    out.nrow[s++] = q;
            }
            std::sort(out.nrow.begin() + out.ncol_ptr[p] + 1, out.nrow.begin() + out.ncol_ptr[p + 1]);
        }
    });
    finalize_layout(out, 1, 0);
    return std::string();
}

void finalize_layout(Symbolic& s, int world, int rank)
{
    const uint32_t F = (uint32_t)s.fronts.size();
    s.world = world;
    s.rank = rank;
    s.cut_level = 1 << 30;
    std::vector<std::vector<uint32_t>> children(F);
    std::vector<double> sub(F, 0.0);
    for (uint32_t f = 0; f < F; ++f) {
        sub[f] += s.fronts[f].work;
        s.fronts[f].owner = 0;
        s.fronts[f].top = 0;
        if (s.fronts[f].parent >= 0) {
            children[s.fronts[f].parent].push_back(f);
            sub[s.fronts[f].parent] += sub[f];   // children precede parents in elimination order
        }
    }
    if (world > 1) {
        std::vector<uint32_t> cand;
        double total = 0;
        for (uint32_t f = 0; f < F; ++f)
            if (s.fronts[f].parent < 0) {
                cand.push_back(f);
                total += sub[f];
            }
        // split the heaviest subtree while there are too few, or it alone exceeds a fair share
        // (GADJ_MG_SUBTREES=n: at least n subtrees — a deeper cut than the rank count needs, for experiments)
        size_t min_subtrees = (size_t)world;
        if (const char* e = getenv("GADJ_MG_SUBTREES"))
            min_subtrees = std::max<size_t>(min_subtrees, (size_t)atoi(e));
        for (;;) {
            size_t best = cand.size();
            for (size_t i = 0; i < cand.size(); ++i)
                if (!children[cand[i]].empty() && (best == cand.size() || sub[cand[i]] > sub[cand[best]]))
                    best = i;
            if (best == cand.size())
                break;
            bool heaviest_is_best = true;
            for (uint32_t c : cand)
                if (sub[c] > sub[cand[best]])
                    heaviest_is_best = false;
            if (cand.size() >= min_subtrees && !(heaviest_is_best && sub[cand[best]] > total / (double)min_subtrees))
                break;
            uint32_t f = cand[best];
            s.fronts[f].top = 1;
            cand.erase(cand.begin() + best);
            for (uint32_t c : children[f])
                cand.push_back(c);
        }
        // longest-processing-time packing of the subtrees
        std::sort(cand.begin(), cand.end(), [&](uint32_t a, uint32_t b) { return sub[a] > sub[b] || (sub[a] == sub[b] && a < b); });
        std::vector<double> load(world, 0.0);
        std::vector<int32_t> root_owner(F, -1);
        for (uint32_t c : cand) {
            int r = (int)(std::min_element(load.begin(), load.end()) - load.begin());
            load[r] += sub[c];
            root_owner[c] = r;
        }
        // propagate subtree ownership downwards (parents have larger indices than children)
        for (uint32_t fi = F; fi-- > 0;) {
            Front& f = s.fronts[fi];
            if (f.top)
                continue;
            if (root_owner[fi] >= 0)
                f.owner = root_owner[fi];
            else
                f.owner = s.fronts[f.parent].owner;
        }
        // A top front is assembled and its pivot panel factorised by one rank.  Level by level (the fronts of a level are
        // factorised side by side): the heaviest front goes to the rank with the least work so far, no rank takes two
        // fronts of a level while another has none.
        s.rank_load.assign(world, 0.0);
        for (uint32_t fi = 0; fi < F; ++fi)
            if (!s.fronts[fi].top)
                s.rank_load[s.fronts[fi].owner] += s.fronts[fi].work;
        {
            std::vector<double> load = s.rank_load;
            int32_t maxlv = 0;
            for (uint32_t fi = 0; fi < F; ++fi)
                if (s.fronts[fi].top) {
                    s.cut_level = std::min(s.cut_level, s.fronts[fi].level);
                    maxlv = std::max(maxlv, s.fronts[fi].level);
                }
            for (int32_t lv = s.cut_level; lv <= maxlv; ++lv) {
                std::vector<uint32_t> fl;
                for (uint32_t fi = 0; fi < F; ++fi)
                    if (s.fronts[fi].top && s.fronts[fi].level == lv)
                        fl.push_back(fi);
                std::sort(fl.begin(), fl.end(), [&](uint32_t a, uint32_t b) {
                    return s.fronts[a].work > s.fronts[b].work || (s.fronts[a].work == s.fronts[b].work && a < b);
                });
                std::vector<int> taken(world, 0);
                for (uint32_t fi : fl) {
                    int best = -1;
                    const int round = *std::min_element(taken.begin(), taken.end());
                    for (int r = 0; r < world; ++r)
                        if (taken[r] == round && (best < 0 || load[r] < load[best]))
                            best = r;
                    Front& f = s.fronts[fi];
                    f.owner = best;
                    taken[best]++;
                    const double k = f.k, r = f.r;
                    load[best] += k * k * k / 3.0 + k * k * r;   // pivot panel + W; the Schur update and the inverse are shared out
                }
            }
        }
    }
    // storage: every top front first, in front order — the same offsets on every rank, so that a tile finished by one rank
    // can be stored into all replicas at (peer base + local offset) — then the fronts of this rank's subtrees
    uint64_t off = 0;
    s.my_factor_flops = s.my_inverse_flops = 0;
    for (int pass = 0; pass < 2; ++pass) {
        for (uint32_t fi = 0; fi < F; ++fi) {
            Front& f = s.fronts[fi];
            if ((pass == 0) != (f.top != 0))
                continue;
            if (f.top || f.owner == rank) {
                f.panel_off = off;
                uint64_t sz = (uint64_t)f.m * f.ldk;
                off += (sz + 15) & ~(uint64_t)15;
            } else
                f.panel_off = NO_DEST;
            double k = f.k, r = f.r;
            const double ff = k * k * k / 3.0 + k * k * r + k * r * r;
            const double fi2 = 2.0 * k * k * k / 3.0 + 2.0 * k * k * r + 2.0 * k * r * r + 2.0 * k * k * r;
            if (f.top) {   // the tiles of a top front are shared out among the ranks
                s.my_factor_flops += ff / world;
                s.my_inverse_flops += fi2 / world;
            } else if (f.owner == rank) {
                s.my_factor_flops += ff;
                s.my_inverse_flops += fi2;
            }
        }
        if (pass == 0)
            s.top_panel_doubles = off;
    }
    s.panel_doubles = off;
    s.pos_owned.assign(s.nstn, 0);
    // destinations of N's blocks: assembled into a front only by its owner
    const uint64_t nslots = s.ncol_ptr[s.nstn];
    s.ndest.assign(nslots, NO_DEST);
    s.ndest_ld.assign(nslots, 0);
    parallel_for(s.nstn, [&](uint64_t p0, uint64_t p1) {
        for (uint64_t p = p0; p < p1; ++p) {
            const Front& fr = s.fronts[s.front_of_pos[p]];
            if (fr.owner != rank)
                continue;
            s.pos_owned[p] = 1;
            const uint32_t own_end = fr.own_begin + fr.own_count;
            const uint32_t* b = s.bnd.data() + fr.bnd_begin;
            const uint64_t col = 3ull * (p - fr.own_begin);
            for (uint64_t t = s.ncol_ptr[p]; t < s.ncol_ptr[p + 1]; ++t) {
                uint32_t q = s.nrow[t];
                uint64_t row;
                if (q < own_end)
                    row = 3ull * (q - fr.own_begin);
                else {
                    const uint32_t* it = std::lower_bound(b, b + fr.bnd_count, q);
                    row = 3ull * (fr.own_count + (uint32_t)(it - b));
                }
                s.ndest[t] = fr.panel_off + row * fr.ldk + col;
                s.ndest_ld[t] = fr.ldk;
            }
        }
    });
}

}  // namespace gadj
