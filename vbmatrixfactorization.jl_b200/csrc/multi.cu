// Single-process multi-device entry points of the C ABI (include/vbmf_b200.h, "several GPUs from ONE host process").
//
// The reference is one Julia process making plain function calls (src/vbmf.jl:175, src/vbmf_sparse.jl:344,
// src/vbmf_dual.jl:455): a drop-in caller cannot be asked to launch one process per GPU.  A multi-device context owns one
// ordinary context per GPU (communicators from ncclCommInitAll), splits the columns of ONE host Y over them, and runs every
// per-device call on its own host thread -- the NCCL all-reduces inside the loop need all ranks in flight at once.  The
// state structs stay the caller's full-size arrays: per-device views are pointer offsets where the layout is contiguous in
// the sharded index (vec(A'), CA, beta, blocks, YHat columns) and small gathered copies where it is not (column-major AHat).
#include "../../include/vbmf_b200.h"
#include "common.cuh"

#include <algorithm>
#include <cstring>
#include <string>
#include <thread>
#include <type_traits>
#include <vector>

namespace vb {
int ctx_create_with_comm(int device, int rank, int world, void* comm, vbmf_b200_ctx** out);
int nccl_comm_init_all(void** comms, int ndev, const int* devs);
}
using namespace vb;

struct vbmf_b200_mctx {
    int ndev = 0;
    std::vector<int> devs;
    std::vector<vbmf_b200_ctx*> ctx;
    int64_t L = 0, M = 0;
    std::vector<int64_t> off, cnt;
    bool have_Y = false;
};

// one host thread per device; the first failing device's message becomes the caller's last_error
template <class F>
static int on_all(vbmf_b200_mctx* m, F&& fn) {
    const int n = m->ndev;
    std::vector<int> rcs(n, 0);
    std::vector<std::string> errs(n);
    auto body = [&](int i) {
        cudaSetDevice(m->devs[i]);
        rcs[i] = fn(i);
        if (rcs[i] != 0) errs[i] = vbmf_b200_last_error();
    };
    if (n == 1) body(0);
    else {
        std::vector<std::thread> th;
        th.reserve(n);
        for (int i = 0; i < n; ++i) th.emplace_back(body, i);
        for (auto& t : th) t.join();
    }
    int rc = 0;
    for (int i = 0; i < n; ++i) if (rcs[i] == -1) { set_error("device %d: %s", m->devs[i], errs[i].c_str()); return -1; }
    for (int i = 0; i < n; ++i) if (rcs[i] != 0) { set_error("device %d: %s", m->devs[i], errs[i].c_str()); rc = rcs[i]; break; }
    return rc;
}

static void split_columns(vbmf_b200_mctx* m, int64_t M) {
    const int n = m->ndev;
    m->off.assign(n, 0); m->cnt.assign(n, 0);
    const int64_t base = M / n, rem = M % n;
    int64_t o = 0;
    for (int i = 0; i < n; ++i) { m->cnt[i] = base + (i < rem ? 1 : 0); m->off[i] = o; o += m->cnt[i]; }
}

extern "C" int vbmf_b200_mctx_create(int ndev, const int* devices, vbmf_b200_mctx** out) {
    if (!out) { set_error("mctx_create: out is NULL"); return -1; }
    *out = nullptr;
    const int avail = vbmf_b200_device_count();
    if (avail == 0) { set_error("no CUDA device available: vbmf_b200 has no CPU fallback"); return -1; }
    if (ndev <= 0) ndev = avail;                       // all visible devices
    if (ndev > avail && devices == nullptr) { set_error("mctx_create: %d devices requested, %d visible", ndev, avail); return -1; }
    vbmf_b200_mctx* m = new vbmf_b200_mctx();
    m->ndev = ndev;
    for (int i = 0; i < ndev; ++i) m->devs.push_back(devices ? devices[i] : i);
    for (int i = 0; i < ndev; ++i)
        for (int k = 0; k < i; ++k)
            if (m->devs[i] == m->devs[k]) { set_error("mctx_create: device %d listed twice", m->devs[i]); delete m; return -1; }
    std::vector<void*> comms(ndev, nullptr);
    if (ndev > 1 && nccl_comm_init_all(comms.data(), ndev, m->devs.data())) { delete m; return -1; }
    m->ctx.assign(ndev, nullptr);
    for (int i = 0; i < ndev; ++i) {
        if (ctx_create_with_comm(m->devs[i], i, ndev, comms[i], &m->ctx[i])) {
            for (int k = 0; k < i; ++k) vbmf_b200_ctx_destroy(m->ctx[k]);
            delete m;
            return -1;
        }
    }
    *out = m;
    return 0;
}

extern "C" int vbmf_b200_mctx_destroy(vbmf_b200_mctx* m) {
    if (!m) return 0;
    for (auto* c : m->ctx) vbmf_b200_ctx_destroy(c);
    delete m;
    return 0;
}

extern "C" int vbmf_b200_mctx_ndev(vbmf_b200_mctx* m) { return m ? m->ndev : 0; }

extern "C" int vbmf_b200_mctx_ctx(vbmf_b200_mctx* m, int i, vbmf_b200_ctx** out) {
    if (!m || !out || i < 0 || i >= m->ndev) { set_error("mctx_ctx: bad argument"); return -1; }
    *out = m->ctx[i];
    return 0;
}

extern "C" int vbmf_b200_mctx_shard(vbmf_b200_mctx* m, int i, int64_t* col_offset, int64_t* n_cols) {
    if (!m || i < 0 || i >= m->ndev || !m->have_Y) { set_error("mctx_shard: bad argument or no Y attached"); return -1; }
    if (col_offset) *col_offset = m->off[i];
    if (n_cols) *n_cols = m->cnt[i];
    return 0;
}

extern "C" int vbmf_b200_mctx_attach_Y(vbmf_b200_mctx* m, const double* Y, int64_t L, int64_t M, int64_t ldY) {
    if (!m || !Y) { set_error("mctx_attach_Y: NULL argument"); return -1; }
    if (ldY < L) { set_error("mctx_attach_Y: ldY < L"); return -1; }
    split_columns(m, M);
    m->have_Y = false;
    const int rc = on_all(m, [&](int i) { return vbmf_b200_attach_Y(m->ctx[i], Y + (size_t)m->off[i] * (size_t)ldY, L, m->cnt[i], ldY, M, m->off[i]); });
    if (rc) return rc;
    m->L = L; m->M = M; m->have_Y = true;
    return 0;
}

extern "C" int vbmf_b200_mctx_synth_Y(vbmf_b200_mctx* m, int64_t L, int64_t M, int rank, double noise, uint64_t seed) {
    if (!m) { set_error("mctx_synth_Y: NULL argument"); return -1; }
    split_columns(m, M);
    m->have_Y = false;
    const int rc = on_all(m, [&](int i) { return vbmf_b200_synth_Y(m->ctx[i], L, m->cnt[i], M, m->off[i], rank, noise, seed); });
    if (rc) return rc;
    m->L = L; m->M = M; m->have_Y = true;
    return 0;
}

extern "C" int vbmf_b200_mctx_trYTY(vbmf_b200_mctx* m, double* out) {
    if (!m || !m->have_Y) { set_error("mctx_trYTY: no Y attached"); return -1; }
    return vbmf_b200_trYTY(m->ctx[0], out);
}

// ---- per-device views of the caller's full-size state -------------------------------------------------------------------
// column-major M x W host matrix <-> the rows [off, off + cnt) as a contiguous cnt x W matrix
static void gather_rows(const double* src, int64_t M, int64_t W, int64_t off, int64_t cnt, std::vector<double>& dst) {
    dst.resize((size_t)std::max<int64_t>(cnt * W, 1));
    if (src) for (int64_t h = 0; h < W; ++h) memcpy(dst.data() + h * cnt, src + h * M + off, (size_t)cnt * 8);
}
static void scatter_rows(double* dst, int64_t M, int64_t W, int64_t off, int64_t cnt, const std::vector<double>& src) {
    if (dst) for (int64_t h = 0; h < W; ++h) memcpy(dst + h * M + off, src.data() + h * cnt, (size_t)cnt * 8);
}
static void split_labels(const int64_t* labels, int64_t n, int64_t off, int64_t cnt, std::vector<int64_t>& out) {
    out.clear();
    for (int64_t k = 0; k < n; ++k) if (labels[k] > off && labels[k] <= off + cnt) out.push_back(labels[k] - off);
}
static int check_full(vbmf_b200_mctx* m, int64_t L, int64_t M) {
    if (!m->have_Y) { set_error("attach Y to the multi-device context first"); return -1; }
    if (L != m->L || M != m->M) { set_error("state is %lld x %lld but the attached Y is %lld x %lld", (long long)L, (long long)M, (long long)m->L, (long long)m->M); return -1; }
    return 0;
}
template <class T> static T* at(T* p, size_t o) { return p ? p + o : nullptr; }

// fields shared by the sparse / dual / trial structs: views for device i; `dl` = download view (replicated outputs only on i = 0)
template <class ST>
static void sparse_like_views(const ST* g, int i, int64_t off, int64_t cnt, std::vector<double>& tmpA, ST& up, ST& dl) {
    const int64_t H = g->H;
    up = *g;
    up.M = cnt; up.MH = cnt * H;
    gather_rows(g->AHat, g->M, H, off, cnt, tmpA);
    up.AHat = tmpA.data();
    up.ATVecHat = at(g->ATVecHat, (size_t)off * H);
    up.diagSigmaATVec = at(g->diagSigmaATVec, (size_t)off * H);
    up.CA = at(g->CA, (size_t)off * H);
    up.beta = at(g->beta, (size_t)off * H);
    up.SigmaATVec_blocks = at(g->SigmaATVec_blocks, (size_t)off * H * H);
    up.YHat = at(g->YHat, (size_t)off * g->L);
    dl = up;
    if (i != 0) {
        dl.SigmaA = nullptr; dl.BHat = nullptr; dl.SigmaB = nullptr; dl.CB = nullptr; dl.delta = nullptr;
        dl.sigmaVecHat = nullptr; dl.etaVec = nullptr; dl.zetaVec = nullptr;
    }
}

extern "C" int vbmf_b200_mctx_dense_run(vbmf_b200_mctx* m, vbmf_b200_dense_state* st, int64_t niter, double eps, int est_covs,
                                        int est_var, int norm_mode, int64_t* iters, double* d) {
    if (!m || !st) { set_error("NULL argument"); return -1; }
    if (check_full(m, st->L, st->M)) return -1;
    std::vector<int64_t> its(m->ndev, 0);
    std::vector<double> ds(m->ndev, 0.0), s2(m->ndev, 0.0);
    const int rc = on_all(m, [&](int i) {
        const int64_t off = m->off[i], cnt = m->cnt[i], H = st->H;
        std::vector<double> tmpA;
        std::vector<int64_t> lab;
        vbmf_b200_dense_state up = *st;
        up.M = cnt;
        gather_rows(st->AHat, st->M, H, off, cnt, tmpA);
        up.AHat = tmpA.data();
        split_labels(st->labels, st->n_labels, off, cnt, lab);
        up.n_labels = (int64_t)lab.size(); up.labels = lab.empty() ? nullptr : lab.data();
        up.YHat = at(st->YHat, (size_t)off * st->L);
        vbmf_b200_solver* s = nullptr;
        if (vbmf_b200_solver_create(m->ctx[i], VBMF_B200_DENSE, H, st->H1, up.n_labels, up.labels, 0, &s)) return -1;
        int r = vbmf_b200_dense_upload(s, &up);
        if (!r) r = vbmf_b200_solver_run(s, niter, eps, (est_covs ? VBMF_B200_EST_COVS : 0) | (est_var ? VBMF_B200_EST_VAR : 0), norm_mode, &its[i], &ds[i]);
        if (r == 0 || r == -2) {
            vbmf_b200_dense_state dl = up;
            if (i != 0) { dl.BHat = nullptr; dl.SigmaA = nullptr; dl.SigmaB = nullptr; dl.CA = nullptr; dl.CB = nullptr; dl.invCA = nullptr; dl.invCB = nullptr; }
            const int r2 = vbmf_b200_dense_download(s, &dl);
            if (r2) r = r2;
            else { scatter_rows(st->AHat, st->M, H, off, cnt, tmpA); s2[i] = dl.sigma2; }
        }
        vbmf_b200_solver_destroy(s);
        return r;
    });
    if (rc == 0 || rc == -2) { st->sigma2 = s2[0]; if (iters) *iters = its[0]; if (d) *d = ds[0]; }
    return rc;
}

template <class ST, class Create, class Upload, class Download, class Extra>
static int run_sparse_like(vbmf_b200_mctx* m, ST* st, int64_t niter, double eps, int flags, int norm_mode, int64_t* iters, double* d,
                           Create create, Upload upload, Download download, Extra extra) {
    if (check_full(m, st->L, st->M)) return -1;
    std::vector<int64_t> its(m->ndev, 0);
    std::vector<double> ds(m->ndev, 0.0);
    std::vector<ST> outs(m->ndev);
    const int rc = on_all(m, [&](int i) {
        const int64_t off = m->off[i], cnt = m->cnt[i];
        std::vector<double> tmpA;
        ST up, dl;
        sparse_like_views(st, i, off, cnt, tmpA, up, dl);
        vbmf_b200_solver* s = nullptr;
        if (create(i, off, cnt, up, dl, &s)) return -1;
        int r = upload(s, &up);
        if (!r) r = vbmf_b200_solver_run(s, niter, eps, flags, norm_mode, &its[i], &ds[i]);
        if (r == 0 || r == -2) {
            const int r2 = download(s, &dl);
            if (r2) r = r2;
            else { scatter_rows(st->AHat, st->M, st->H, off, cnt, tmpA); extra(i, off, cnt, dl); outs[i] = dl; }
        }
        vbmf_b200_solver_destroy(s);
        return r;
    });
    if (rc == 0 || rc == -2) {
        st->sigmaHat = outs[0].sigmaHat; st->zeta = outs[0].zeta; st->eta = outs[0].eta;
        if (iters) *iters = its[0];
        if (d) *d = ds[0];
    }
    return rc;
}

extern "C" int vbmf_b200_mctx_sparse_run(vbmf_b200_mctx* m, vbmf_b200_sparse_state* st, int64_t niter, double eps, int diag_var,
                                         int full_cov, int est_cb, int norm_mode, int64_t* iters, double* d) {
    if (!m || !st) { set_error("NULL argument"); return -1; }
    const int flags = (diag_var ? VBMF_B200_DIAG_VAR : 0) | (full_cov ? VBMF_B200_FULL_COV : 0) | (est_cb ? VBMF_B200_EST_CB : 0);
    std::vector<std::vector<int64_t>> labs(m->ndev);
    return run_sparse_like(m, st, niter, eps, flags, norm_mode, iters, d,
        [&](int i, int64_t off, int64_t cnt, vbmf_b200_sparse_state& up, vbmf_b200_sparse_state& dl, vbmf_b200_solver** s) {
            split_labels(st->labels, st->n_labels, off, cnt, labs[i]);
            up.n_labels = dl.n_labels = (int64_t)labs[i].size();
            up.labels = dl.labels = labs[i].empty() ? nullptr : labs[i].data();
            return vbmf_b200_solver_create(m->ctx[i], VBMF_B200_SPARSE, st->H, st->H1, up.n_labels, up.labels, st->SigmaATVec_blocks != nullptr, s);
        },
        vbmf_b200_sparse_upload, vbmf_b200_sparse_download, [](int, int64_t, int64_t, vbmf_b200_sparse_state&) {});
}

extern "C" int vbmf_b200_mctx_dual_run(vbmf_b200_mctx* m, vbmf_b200_dual_state* st, int64_t niter, double eps, int diag_var,
                                       int full_cov, int est_priors, int est_cb, int norm_mode, int64_t* iters, double* d) {
    if (!m || !st) { set_error("NULL argument"); return -1; }
    if (st->H < st->H0) { set_error("H must be at least H0!"); return -1; }
    const int flags = (diag_var ? VBMF_B200_DIAG_VAR : 0) | (full_cov ? VBMF_B200_FULL_COV : 0) | (est_cb ? VBMF_B200_EST_CB : 0) |
                      (est_priors ? VBMF_B200_EST_PRIORS : 0);
    const int64_t H0 = st->H0, H1 = st->H - st->H0;
    std::vector<std::vector<double>> t0(m->ndev), t1(m->ndev);
    const int rc = run_sparse_like(m, st, niter, eps, flags, norm_mode, iters, d,
        [&](int i, int64_t off, int64_t cnt, vbmf_b200_dual_state& up, vbmf_b200_dual_state& dl, vbmf_b200_solver** s) {
            t0[i].assign((size_t)std::max<int64_t>(cnt * H0, 1), 0.0); t1[i].assign((size_t)std::max<int64_t>(cnt * H1, 1), 0.0);
            for (vbmf_b200_dual_state* v : {&up, &dl}) {
                v->A0Hat = st->A0Hat ? t0[i].data() : nullptr; v->A1Hat = st->A1Hat ? t1[i].data() : nullptr;
                v->CA0 = at(st->CA0, (size_t)off * H0); v->beta0 = at(st->beta0, (size_t)off * H0);
                v->CA1 = at(st->CA1, (size_t)off * H1); v->beta1 = at(st->beta1, (size_t)off * H1);
            }
            if (i != 0) dl.alpha = nullptr;
            return vbmf_b200_solver_create(m->ctx[i], VBMF_B200_DUAL, st->H, H0, 0, nullptr, st->SigmaATVec_blocks != nullptr, s);
        },
        vbmf_b200_dual_upload, vbmf_b200_dual_download,
        [&](int i, int64_t off, int64_t cnt, vbmf_b200_dual_state& dl) {
            scatter_rows(st->A0Hat, st->M, H0, off, cnt, t0[i]); scatter_rows(st->A1Hat, st->M, H1, off, cnt, t1[i]);
            if (i == 0) { st->alpha00 = dl.alpha00; st->beta00 = dl.beta00; st->alpha01 = dl.alpha01; st->beta01 = dl.beta01; st->alpha0 = dl.alpha0; st->alpha1 = dl.alpha1; }
        });
    return rc;
}

extern "C" int vbmf_b200_mctx_trial_run(vbmf_b200_mctx* m, vbmf_b200_trial_state* st, int64_t niter, double eps, int diag_var,
                                        int full_cov, int est_priors, int est_cb, int norm_mode, int64_t* iters, double* d) {
    if (!m || !st) { set_error("NULL argument"); return -1; }
    if (st->H < st->H0) { set_error("H must be at least H0!"); return -1; }
    const int flags = (diag_var ? VBMF_B200_DIAG_VAR : 0) | (full_cov ? VBMF_B200_FULL_COV : 0) | (est_cb ? VBMF_B200_EST_CB : 0) |
                      (est_priors ? VBMF_B200_EST_PRIORS : 0);
    const int64_t H0 = st->H0, H1 = st->H - st->H0, M0 = st->M0, M1 = st->M - st->M0;
    std::vector<std::vector<double>> t1(m->ndev), t2(m->ndev), t3(m->ndev);
    auto rows2 = [&](int64_t off, int64_t cnt, int64_t* o2, int64_t* n2, int64_t* o3, int64_t* n3) {
        const int64_t a = std::min(off, M0), b = std::min(off + cnt, M0);      // rows of this shard in group 2: [a, b)
        *o2 = a; *n2 = b - a; *o3 = std::max<int64_t>(off - M0, 0); *n3 = cnt - (b - a);
    };
    return run_sparse_like(m, st, niter, eps, flags, norm_mode, iters, d,
        [&](int i, int64_t off, int64_t cnt, vbmf_b200_trial_state& up, vbmf_b200_trial_state& dl, vbmf_b200_solver** s) {
            int64_t o2, n2, o3, n3;
            rows2(off, cnt, &o2, &n2, &o3, &n3);
            t1[i].assign((size_t)std::max<int64_t>(cnt * H0, 1), 0.0); t2[i].assign((size_t)std::max<int64_t>(n2 * H1, 1), 0.0); t3[i].assign((size_t)std::max<int64_t>(n3 * H1, 1), 0.0);
            for (vbmf_b200_trial_state* v : {&up, &dl}) {
                v->A1Hat = st->A1Hat ? t1[i].data() : nullptr; v->A2Hat = st->A2Hat ? t2[i].data() : nullptr; v->A3Hat = st->A3Hat ? t3[i].data() : nullptr;
                v->CA1 = at(st->CA1, (size_t)off * H0); v->beta1 = at(st->beta1, (size_t)off * H0);
                v->CA2 = at(st->CA2, (size_t)o2 * H1); v->beta2 = at(st->beta2, (size_t)o2 * H1);
                v->CA3 = at(st->CA3, (size_t)o3 * H1); v->beta3 = at(st->beta3, (size_t)o3 * H1);
            }
            if (i != 0) dl.alpha = nullptr;
            return vbmf_b200_solver_create_trial(m->ctx[i], st->H, H0, M0, st->SigmaATVec_blocks != nullptr, s);
        },
        vbmf_b200_trial_upload, vbmf_b200_trial_download,
        [&](int i, int64_t off, int64_t cnt, vbmf_b200_trial_state& dl) {
            int64_t o2, n2, o3, n3;
            rows2(off, cnt, &o2, &n2, &o3, &n3);
            scatter_rows(st->A1Hat, st->M, H0, off, cnt, t1[i]);
            scatter_rows(st->A2Hat, M0, H1, o2, n2, t2[i]);
            scatter_rows(st->A3Hat, M1, H1, o3, n3, t3[i]);
            if (i == 0) {
                st->alpha01 = dl.alpha01; st->beta01 = dl.beta01; st->alpha02 = dl.alpha02; st->beta02 = dl.beta02; st->alpha03 = dl.alpha03; st->beta03 = dl.beta03;
                st->alpha1 = dl.alpha1; st->alpha2 = dl.alpha2; st->alpha3 = dl.alpha3;
            }
        });
}

// lowerBound / lowerBoundTrimmed on a full-size sparse / dual / trial state (src/vbmf_sparse.jl:435-489, src/vbmf_dual.jl:556-617)
extern "C" int vbmf_b200_mctx_lower_bound(vbmf_b200_mctx* m, int kind, void* state, double trim, int trimmed, double* out) {
    if (!m || !state || !out) { set_error("NULL argument"); return -1; }
    std::vector<double> lbs(m->ndev, 0.0);
    int rc = -1;
    auto go = [&](auto* st, auto create, auto upload) {
        if (check_full(m, st->L, st->M)) return -1;
        return on_all(m, [&](int i) {
            typename std::remove_pointer<decltype(st)>::type up, dl;
            std::vector<double> tmpA;
            sparse_like_views(st, i, m->off[i], m->cnt[i], tmpA, up, dl);
            vbmf_b200_solver* s = nullptr;
            if (create(i, up, &s)) return -1;
            int r = upload(s, &up);
            if (!r) r = vbmf_b200_solver_lower_bound(s, trim, trimmed, &lbs[i]);
            vbmf_b200_solver_destroy(s);
            return r;
        });
    };
    if (kind == VBMF_B200_SPARSE) {
        auto* st = (vbmf_b200_sparse_state*)state;
        rc = go(st, [&](int i, vbmf_b200_sparse_state& up, vbmf_b200_solver** s) {
            up.n_labels = 0; up.labels = nullptr;       // lowerBound does not mask
            return vbmf_b200_solver_create(m->ctx[i], VBMF_B200_SPARSE, st->H, st->H1, 0, nullptr, 0, s); }, vbmf_b200_sparse_upload);
    } else if (kind == VBMF_B200_DUAL) {
        auto* st = (vbmf_b200_dual_state*)state;
        rc = go(st, [&](int i, vbmf_b200_dual_state&, vbmf_b200_solver** s) { return vbmf_b200_solver_create(m->ctx[i], VBMF_B200_DUAL, st->H, st->H0, 0, nullptr, 0, s); },
                vbmf_b200_dual_upload);
    } else if (kind == VBMF_B200_TRIAL) {
        auto* st = (vbmf_b200_trial_state*)state;
        rc = go(st, [&](int i, vbmf_b200_trial_state&, vbmf_b200_solver** s) { return vbmf_b200_solver_create_trial(m->ctx[i], st->H, st->H0, st->M0, 0, s); },
                vbmf_b200_trial_upload);
    } else { set_error("lowerBound exists for sparse / dual / trial parameters only"); return -1; }
    if (rc == 0) *out = lbs[0];
    return rc;
}
