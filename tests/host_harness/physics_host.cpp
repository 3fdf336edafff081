// tests/host_harness/physics_host.cpp — TEST-ONLY build of the device physics headers with g++.
//
// Compiles physicssimulator_b200/csrc/ps_physics.cuh (integration, narrow phase, per-pair solve) for
// the host so that `-m "not gpu"` tests can compare those routines with the oracle bit for bit in a
// container without a GPU.  It walks ALL pairs serially like the kernel's exact path.  It is not part
// of libps_b200.so and nothing in the product links or loads it.
#include <cmath>
#include <cstdint>
#include <vector>

#include "../../physicssimulator_b200/csrc/ps_physics.cuh"
#include "../../physicssimulator_b200/csrc/ps_octree.hpp"
#include "../../physicssimulator_b200/csrc/ps_coop.cuh"
#include <thread>

#include <cstring>
#include <map>

using namespace psp;

namespace {
void ld_dyn(const double* a, int N, int b, Dyn& d)
{
    d.x = mk(a[b], a[N + b], a[2 * N + b]);
    d.v = mk(a[3 * N + b], a[4 * N + b], a[5 * N + b]);
    for (int k = 0; k < 9; k++) d.R.a[k] = a[(6 + k) * N + b];
    d.L = mk(a[15 * N + b], a[16 * N + b], a[17 * N + b]);
}
void st_dyn(double* a, int N, int b, const Dyn& d)
{
    a[b] = d.x.x; a[N + b] = d.x.y; a[2 * N + b] = d.x.z;
    a[3 * N + b] = d.v.x; a[4 * N + b] = d.v.y; a[5 * N + b] = d.v.z;
    for (int k = 0; k < 9; k++) a[(6 + k) * N + b] = d.R.a[k];
    a[15 * N + b] = d.L.x; a[16 * N + b] = d.L.y; a[17 * N + b] = d.L.z;
}
void ld_par(const double* a, int N, int b, Par& p)
{
    p.inv_mass = a[b];
    p.is_static = (p.inv_mass == 0.0);
    p.iinv[0] = a[N + b]; p.iinv[1] = a[2 * N + b]; p.iinv[2] = a[3 * N + b];
    p.F = mk(a[4 * N + b], a[5 * N + b], a[6 * N + b]);
    p.T = mk(a[7 * N + b], a[8 * N + b], a[9 * N + b]);
    p.e = a[10 * N + b];
    p.mu = a[11 * N + b];
    p.dims[0] = a[12 * N + b]; p.dims[1] = a[13 * N + b]; p.dims[2] = a[14 * N + b];
    p.shape = (int)a[15 * N + b];
}
}  // namespace

// box-box narrow phase by a team of 8 host threads (ps_coop.cuh under the barrier emulation of ps_team.cuh)
static int team_detect_bb(const Dyn& di, const Par& pi, const Dyn& dj, const Par& pj, double eps, Contact* cs, int& err, bool& need_literal)
{
    using namespace psc;
    static pstm::TeamShared<kTeam> sh;
    static TeamScratch ts;
    TeamContacts outs[kTeam];
    int ncs[kTeam], errs[kTeam];
    bool nl[kTeam];
    std::thread th[kTeam];
    for (int l = 0; l < kTeam; l++)
        th[l] = std::thread([&, l] {
            pstm::Team<kTeam> tm(&sh, l);
            errs[l] = 0; nl[l] = false;
            ncs[l] = detect_bb_team(tm, di, pi, dj, pj, eps, ts, outs[l], errs[l], nl[l]);
        });
    for (int l = 0; l < kTeam; l++) th[l].join();
    for (int l = 1; l < kTeam; l++)
        if (ncs[l] != ncs[0] || nl[l] != nl[0] || errs[l] != errs[0]) return -1000;  // lanes must agree
    err = errs[0];
    need_literal = nl[0];
    int seen = 0;
    for (int l = 0; l < kTeam; l++)
        for (int t = 0; t < 2; t++)
            if (outs[l].ord[t] >= 0) { cs[outs[l].ord[t]] = outs[l].c[t]; seen++; }
    if (!need_literal && seen != ncs[0]) return -1001;
    return ncs[0];
}
static bool same_contact(const Contact& a, const Contact& b)
{
    return memcmp(&a.n, &b.n, sizeof(d3)) == 0 && memcmp(&a.p, &b.p, sizeof(d3)) == 0 && memcmp(&a.pen, &b.pen, 8) == 0 && a.feat == b.feat;
}

extern "C" {

// planar layouts as in ps_step_kernel.cuh: dyn[f*N+b] (18 planes), par[f*N+b] (16 planes, plane 0 = 1/mass)
// fast bit 0: the inlined/unrolled forms of the narrow phase (the ones the batch-wide kernels call)
// fast bit 1: the straight-line (select-based) integration (integrate_a)
// fast bit 3: the straight-line solve (resolve_s) instead of the literal one
// fast bit 2: every box-box pair is ALSO run through the team form (ps_coop.cuh, 8 host threads per pair) and compared:
//             counters[5] team runs, [6] literal requested, [7] mismatches
void hh_world_step_v(int N, const double* par, double* dyn, const ps_step_config* cfg, int nsteps, int64_t* counters /*tested,hits,contacts,exc,overflow*/,
                     int32_t* last_pairs, uint32_t* last_feats, double* last_contacts /*7 per contact*/, int32_t* last_counts /*pairs,contacts*/, int cap_pairs, int cap_contacts,
                     int fast)
{
    StepParams sp;
    sp.eps = cfg->epsilon; sp.baumgarte = cfg->baumgarte; sp.slop = cfg->allowed_penetration; sp.max_corr = cfg->max_correction_velocity;
    sp.dt = cfg->dt; sp.iterations = cfg->solver_iterations; sp.friction = cfg->enable_friction; sp.integrator = cfg->integrator_kind;
    std::vector<double> pred(18 * (size_t)N);
    std::vector<char> pvalid(N);
    for (int s = 0; s < nsteps; s++) {
        std::fill(pvalid.begin(), pvalid.end(), 0);
        int np = 0, ncs = 0;
        auto get_pred = [&](int b, const Par& p, Dyn& out) {
            if (pvalid[b]) { ld_dyn(pred.data(), N, b, out); return; }
            Dyn cur; ld_dyn(dyn, N, b, cur);
            if (fast & 2) integrate_a(cur, p, sp, out); else integrate(cur, p, sp, out);
            st_dyn(pred.data(), N, b, out);
            pvalid[b] = 1;
        };
        for (int i = 0; i < N; i++)
            for (int j = i + 1; j < N; j++) {
                Par pi, pj; ld_par(par, N, i, pi); ld_par(par, N, j, pj);
                if (pi.is_static && pj.is_static) continue;
                Dyn di, dj; get_pred(i, pi, di); get_pred(j, pj, dj);
                static PairScratch sc;
                Contact* cs = sc.cs;
                int err = 0, ovf = 0;
                counters[0]++;
                int nc = (fast & 1) ? detect_t<true>(di, pi, dj, pj, sp.eps, cs, err, ovf, sc) : detect(di, pi, dj, pj, sp.eps, cs, err, ovf, sc);
                if ((fast & 4) && pi.shape == 1 && pj.shape == 1) {  // the team form must give the literal answer (or ask for it)
                    Contact tc[PS_MAX_CONTACTS];
                    int terr = 0;
                    bool nl = false;
                    int tn = team_detect_bb(di, pi, dj, pj, sp.eps, tc, terr, nl);
                    counters[5]++;
                    if (nl) counters[6]++;
                    else {
                        bool okc = tn == nc && terr == err;
                        for (int k = 0; okc && k < nc; k++) okc = same_contact(tc[k], cs[k]);
                        if (!okc) counters[7]++;
                    }
                }
                counters[3] += err; counters[4] += ovf;
                if (nc <= 0) continue;
                counters[1]++; counters[2] += nc;
                if (s == nsteps - 1 && last_pairs) {
                    if (np < cap_pairs) { last_pairs[3 * np] = i; last_pairs[3 * np + 1] = j; last_pairs[3 * np + 2] = nc; }
                    np++;
                    for (int k = 0; k < nc; k++) {
                        if (ncs < cap_contacts) {
                            double* o = last_contacts + 7 * (size_t)ncs;
                            o[0] = cs[k].n.x; o[1] = cs[k].n.y; o[2] = cs[k].n.z; o[3] = cs[k].p.x; o[4] = cs[k].p.y; o[5] = cs[k].p.z; o[6] = cs[k].pen;
                            last_feats[ncs] = cs[k].feat;
                        }
                        ncs++;
                    }
                }
                if (fast & 8) {  // the straight-line solve of the batch-wide kernels (resolve_s), all its template forms
                    CPFast cpf[PS_MAX_CONTACTS];
                    const bool s1 = pi.is_static && !pj.is_static && bits_of(di.v.x) == 0 && bits_of(di.v.y) == 0 && bits_of(di.v.z) == 0;
                    bool ok;
                    const bool one = nc == 1 && (pi.shape == 0 || pj.shape == 0);
                    if (one) ok = s1 ? resolve_s<1, true>(di, pi, dj, pj, cs, nc, cpf, sp) : resolve_s<1, false>(di, pi, dj, pj, cs, nc, cpf, sp);
                    else ok = s1 ? resolve_s<0, true>(di, pi, dj, pj, cs, nc, cpf, sp) : resolve_s<0, false>(di, pi, dj, pj, cs, nc, cpf, sp);
                    if (!ok) counters[6] += 1000000;  // (cannot happen on the host: plain divisions, finite states)
                } else
                    resolve(di, pi, dj, pj, cs, nc, sc.cp, sp);
                st_dyn(dyn, N, i, di); pvalid[i] = 0;
                st_dyn(dyn, N, j, dj); pvalid[j] = 0;
            }
        for (int b = 0; b < N; b++) {
            Dyn out;
            if (pvalid[b]) ld_dyn(pred.data(), N, b, out);
            else { Par p; ld_par(par, N, b, p); Dyn cur; ld_dyn(dyn, N, b, cur); if (fast & 2) integrate_a(cur, p, sp, out); else integrate(cur, p, sp, out); }
            st_dyn(dyn, N, b, out);
        }
        if (last_counts) { last_counts[0] = np; last_counts[1] = ncs; }
    }
}
void hh_world_step(int N, const double* par, double* dyn, const ps_step_config* cfg, int nsteps, int64_t* counters,
                   int32_t* last_pairs, uint32_t* last_feats, double* last_contacts, int32_t* last_counts, int cap_pairs, int cap_contacts)
{
    hh_world_step_v(N, par, dyn, cfg, nsteps, counters, last_pairs, last_feats, last_contacts, last_counts, cap_pairs, cap_contacts, 0);
}

// ---- the product's host-side SpatialTree (ps_octree.hpp) behind the same hooks as the oracle's tree, so that the
// reference's known-answer facts (PhysicsSimulator.Tests/UtilitiesTests.fs) run against it too ------------------
}  // extern "C"
struct HHTree {
    int ndim;
    pst::Tree<1> t1; pst::Tree<2> t2; pst::Tree<3> t3;
    std::map<int, std::vector<double>> ext;  // id -> (size, pos) per dimension
};
template <int D> static pst::Ext<D> hh_ext(const HHTree* h, int id)
{
    pst::Ext<D> e{};
    auto it = h->ext.find(id);
    for (int k = 0; k < D; k++) { e.size[k] = it->second[2 * k]; e.pos[k] = it->second[2 * k + 1]; }
    return e;
}
extern "C" {
void* hh_tree_create(int max_leaf, int max_depth, int ndim, const double* bounds /* (lo, hi) per dimension */)
{
    if (ndim < 1 || ndim > 3) return nullptr;
    HHTree* h = new HHTree();
    h->ndim = ndim;
    double lo[3], hi[3];
    for (int k = 0; k < ndim; k++) { lo[k] = bounds[2 * k]; hi[k] = bounds[2 * k + 1]; }
    bool ok = ndim == 1 ? h->t1.init(max_leaf, max_depth, lo, hi) : ndim == 2 ? h->t2.init(max_leaf, max_depth, lo, hi) : h->t3.init(max_leaf, max_depth, lo, hi);
    if (!ok) { delete h; return nullptr; }
    return h;
}
void hh_tree_destroy(void* p) { delete (HHTree*)p; }
void hh_tree_set_extent(void* p, int id, const double* e) { HHTree* h = (HHTree*)p; h->ext[id].assign(e, e + 2 * h->ndim); }
void hh_tree_set_root_leaf(void* p, int n, const int32_t* ids)
{
    HHTree* h = (HHTree*)p;
    std::vector<int> v(ids, ids + n);
    if (h->ndim == 1) h->t1.set_root_leaf(v); else if (h->ndim == 2) h->t2.set_root_leaf(v); else h->t3.set_root_leaf(v);
}
void hh_tree_set_root_children(void* p, int nkids, const int32_t* counts, const int32_t* ids)
{
    HHTree* h = (HHTree*)p;
    std::vector<std::vector<int>> kids;
    for (int c = 0, o = 0; c < nkids; o += counts[c], c++) kids.emplace_back(ids + o, ids + o + counts[c]);
    if (h->ndim == 1) h->t1.set_root_children(kids); else if (h->ndim == 2) h->t2.set_root_children(kids); else h->t3.set_root_children(kids);
}
int hh_tree_insert(void* p, int id)
{
    HHTree* h = (HHTree*)p;
    if (!h->ext.count(id)) return -2;
    bool ok = h->ndim == 1 ? h->t1.insert(id, [&](int o) { return hh_ext<1>(h, o); })
            : h->ndim == 2 ? h->t2.insert(id, [&](int o) { return hh_ext<2>(h, o); })
                           : h->t3.insert(id, [&](int o) { return hh_ext<3>(h, o); });
    return ok ? 0 : -1;
}
void hh_tree_remove(void* p, int id)
{
    HHTree* h = (HHTree*)p;
    auto gone = [&](int o) { return o == id; };
    if (h->ndim == 1) h->t1.remove_if(gone); else if (h->ndim == 2) h->t2.remove_if(gone); else h->t3.remove_if(gone);
}
int hh_tree_dump(void* p, char* buf, int cap)
{
    HHTree* h = (HHTree*)p;
    std::string s = h->ndim == 1 ? h->t1.dump() : h->ndim == 2 ? h->t2.dump() : h->t3.dump();
    if ((int)s.size() + 1 > cap) return -1;
    std::memcpy(buf, s.c_str(), s.size() + 1);
    return (int)s.size();
}
int hh_tree_num_buckets(void* p)
{
    HHTree* h = (HHTree*)p;
    int n = 0;
    auto f = [&](const std::vector<int>&) { n++; };
    if (h->ndim == 1) h->t1.for_each_bucket(f); else if (h->ndim == 2) h->t2.for_each_bucket(f); else h->t3.for_each_bucket(f);
    return n;
}
// candidate list of a 3-D tree (BroadPhase.getCollisionCandidates + canIntersect), ids i | j << 16
int hh_tree_candidates(void* p, const unsigned char* is_static, uint32_t* out, int cap)
{
    HHTree* h = (HHTree*)p;
    if (h->ndim != 3) return -1;
    std::vector<uint32_t> v;
    pst::candidates(h->t3, is_static, v);
    for (int k = 0; k < (int)v.size() && k < cap; k++) out[k] = v[k];
    return (int)v.size();
}
// ps_sincos.h (shared by the device path and the oracle) on an array: tests/test_sincos_ulp.py holds it against libm
void hh_sincos(int n, const double* x, double* s, double* c)
{
    for (int k = 0; k < n; k++) ps_sincos(x[k], &s[k], &c[k]);
}
}
