// fast_host.inl -- host-side packing of the constant block consumed by k_step_fast
// (included at the end of engine.cu).  Everything is expressed in block-SORTED
// coordinates j (x_sorted[j] = x[i_of_j[j]]), padded to DP = 8*NT, and the two matrices
// are laid out in m8n8k4 B-fragment order (see warp_matvec8).
static void pack_frag(std::vector<double> &out, size_t off, const std::vector<double> &Mat,
                      int DP, int NT, bool tri) {
    for (int nt = 0; nt < NT; ++nt)
        for (int m = 0; m < NT; ++m) {
            if (tri && m > nt) continue;
            const size_t blk = tri ? (size_t)(nt * (nt + 1)) / 2 + m : (size_t)nt * NT + m;
            for (int lane = 0; lane < 32; ++lane) {
                const int q = lane >> 2, r = lane & 3;
                const int row = 8 * nt + q, col = 8 * m + 2 * r;
                out[off + blk * 64 + lane * 2 + 0] = Mat[(size_t)row * DP + col];
                out[off + blk * 64 + lane * 2 + 1] = Mat[(size_t)row * DP + col + 1];
            }
        }
}

static int pack_fast(cb2_engine *h) {
    h->fast_ready = false;
    h->drag_ready = false;
    h->fastG_ready = false;
    // one likelihood over all D <= 64 parameters: a Gaussian mixture (<= 4 modes, no derived
    // parameters) for the Metropolis and dragging kernels, or the built-in Rosenbrock over
    // the parameters in sampler order for the dragging kernel (kernels_drag.cuh)
    if (h->D > 64 || h->likes.size() != 1) return 0;
    if (!h->exts.empty()) return 0;  // external functions / priors: kernels_ext.cuh only
    const bool gauss = fast_step_supported_like(h->M);
    const bool rosen = h->likes[0].d.kind == 1 && h->drag && h->likes[0].d.dim == h->D;
    if (!gauss && !rosen) return 0;
    const int D = h->D, NT = (D + 7) / 8, DP = 8 * NT;
    const LikeHost &L = h->likes[0];
    const int nm = gauss ? L.d.n_modes : 0;
    std::vector<int> j_of_i(D), ilike_of_i(D, -1);
    for (int j = 0; j < D; ++j) j_of_i[h->i_of_j[j]] = j;
    for (int a = 0; a < D; ++a) {
        if (ilike_of_i[L.idx[a]] != -1) return 0;  // repeated input parameter: general path
        ilike_of_i[L.idx[a]] = a;
    }
    bool tri = true;
    for (int j = 0; j < D; ++j)
        if (ilike_of_i[h->i_of_j[j]] != j) tri = false;
    if (rosen && !tri) return 0;  // Rosenbrock couples neighbours: sorted order = input order
    FastPackDesc P;
    memset(&P, 0, sizeof(P));
    P.NT = NT;
    P.n_modes = nm;
    P.tri_like = tri ? 1 : 0;
    const int blocks_T = NT * (NT + 1) / 2;
    P.blocks_A = tri ? blocks_T : NT * NT;
    int o = 0;
    auto take = [&](int n) { int r = o; o += (n + 1) & ~1; return r; };
    P.off_T = take(blocks_T * 64);
    P.off_A = take(nm * P.blocks_A * 64);
    P.off_mu = take(nm * DP);
    P.off_c0 = take(nm);
    P.off_w = take(nm);
    P.off_lower = take(DP); P.off_upper = take(DP); P.off_loc = take(DP);
    P.off_mls = take(DP); P.off_isc = take(DP); P.off_pa = take(DP); P.off_pb = take(DP);
    P.off_flags = take(DP); P.off_iofj = take(DP);
    P.off_klo = take(DP); P.off_kup = take(DP);
    P.total = o;
    {
        bool ident = (D == DP) && (row_width(h) % 2 == 0);
        for (int j = 0; j < D; ++j) ident = ident && (h->i_of_j[j] == j);
        P.iofj_identity = ident ? 1 : 0;
        bool vec = true;
        for (int b = 0; b < h->n_blocks; ++b)
            if (h->bsize[b] >= 2 && ((h->bsize[b] | h->jstart[b]) & 1)) vec = false;
        P.vec_ok = vec ? 1 : 0;
    }
    if ((size_t)P.total * 8 > 190 * 1024) return 0;  // does not fit: general path
    std::vector<double> pk(P.total, 0.0);
    // T (sorted coordinates already)
    std::vector<double> Mat((size_t)DP * DP, 0.0);
    for (int j = 0; j < D; ++j)
        for (int k = 0; k <= j; ++k) Mat[(size_t)j * DP + k] = h->Trow[(size_t)j * D + k];
    pack_frag(pk, P.off_T, Mat, DP, NT, true);
    for (int km = 0; km < nm; ++km) {
        std::fill(Mat.begin(), Mat.end(), 0.0);
        for (int a = 0; a < D; ++a)
            for (int j = 0; j < D; ++j) {
                const int il = ilike_of_i[h->i_of_j[j]];
                // Linv[a][il] = linvT[il*dim + a]
                Mat[(size_t)a * DP + j] = L.linvT[(size_t)km * D * D + (size_t)il * D + a];
            }
        pack_frag(pk, P.off_A + (size_t)km * P.blocks_A * 64, Mat, DP, NT, tri);
        for (int j = 0; j < D; ++j)
            pk[P.off_mu + km * DP + j] = L.means[(size_t)km * D + ilike_of_i[h->i_of_j[j]]];
        pk[P.off_c0 + km] = L.c0[km];
        pk[P.off_w + km] = L.w[km];
    }
    // flags / i_of_j are stored as int32 inside the (double) pack
    // order-preserving integer keys of the bounds (k_step_pc2 tests them on the integer pipe)
    auto dkey = [](double d) {
        int64_t b;
        memcpy(&b, &d, 8);
        b ^= (b >> 63) & 0x7fffffffffffffffLL;
        double out;
        memcpy(&out, &b, 8);
        return out;
    };
    int32_t *iflags = reinterpret_cast<int32_t *>(pk.data() + P.off_flags);
    int32_t *iiofj = reinterpret_cast<int32_t *>(pk.data() + P.off_iofj);
    for (int j = 0; j < DP; ++j) {
        if (j < D) {
            const int i = h->i_of_j[j];
            pk[P.off_lower + j] = h->lower[i];
            pk[P.off_upper + j] = h->upper[i];
            pk[P.off_klo + j] = dkey(h->lower[i]);
            pk[P.off_kup + j] = dkey(h->upper[i]);
            pk[P.off_loc + j] = h->loc[i];
            pk[P.off_isc + j] = h->pscale[i];
            const int kd = h->prior_kind[i];
            pk[P.off_mls + j] = (kd == 1) ? (-std::log(h->pscale[i]) - CB2_LOG_2PI / 2)
                                : (kd >= 2 ? h->pcn[i] : 0.0);
            pk[P.off_pa + j] = kd >= 2 ? h->pa[i] : 0.0;
            pk[P.off_pb + j] = kd >= 2 ? h->pb[i] : 0.0;
            iflags[j] = (kd != 0 ? 1 : 0) | (h->periodic[i] ? 2 : 0) | (kd << 8);
            iiofj[j] = i;
        } else {
            pk[P.off_lower + j] = -INFINITY;
            pk[P.off_upper + j] = INFINITY;
            pk[P.off_klo + j] = dkey(-INFINITY);
            pk[P.off_kup + j] = dkey(INFINITY);
            pk[P.off_isc + j] = 1.0;
            iflags[j] = 0;
            iiofj[j] = -1;
        }
    }
    int rc = upload(h, h->d_fastpack, pk);
    if (rc) return rc;
    h->drag_extra.like_kind = rosen ? 1 : 0;
    h->drag_extra.like_dim = L.d.dim;
    h->drag_extra.like_scale = L.d.scale;
    // G = (L^-1 P) T for k_step_pc2 (one mode, triangular likelihood matrix): the image of a
    // direction in whitened coordinates without going through delta.  Lower triangular as
    // the product of two lower-triangular matrices.
    if (tri && nm == 1 && !h->drag) {
        std::vector<double> Am((size_t)DP * DP, 0.0), Tm((size_t)DP * DP, 0.0),
            Gm((size_t)DP * DP, 0.0);
        for (int a = 0; a < D; ++a)
            for (int j = 0; j <= a; ++j)
                Am[(size_t)a * DP + j] = L.linvT[(size_t)ilike_of_i[h->i_of_j[j]] * D + a];
        for (int j = 0; j < D; ++j)
            for (int k = 0; k <= j; ++k) Tm[(size_t)j * DP + k] = h->Trow[(size_t)j * D + k];
        for (int a = 0; a < D; ++a)
            for (int k = 0; k <= a; ++k) {
                // one fused multiply-add chain in j order: k_ckpt_pack (kernels_ckpt.cuh)
                // forms the same matrix on the device, bit for bit
                double acc = 0.0;
                for (int j = k; j <= a; ++j)
                    acc = std::fma(Am[(size_t)a * DP + j], Tm[(size_t)j * DP + k], acc);
                Gm[(size_t)a * DP + k] = acc;
            }
        std::vector<double> gk((size_t)blocks_T * 64, 0.0);
        pack_frag(gk, 0, Gm, DP, NT, true);
        if ((rc = upload(h, h->d_fastG, gk))) return rc;
        h->fastG_ready = true;
    }
    h->fast_desc = P;
    h->fast_ready = gauss && !h->drag;
    h->drag_ready = drag_step_supported(h->M, NT);
    return 0;
}

// ---------------------------------------------------------------------------------------
// pack_stream: constant block of the streamed path (kernels_stream.cuh) for 64 < D <= 128.
// Same conventions as pack_fast (block-sorted coordinates, DP = 8 NT, B-fragment order); it
// lives in global memory (the kernels read it through L1/L2), so there is no size limit.
// ---------------------------------------------------------------------------------------
static int pack_stream(cb2_engine *h) {
    h->stream_ready = false;
    // D <= 64 is the territory of the register-resident kernels; the streamed path takes what
    // they refuse (several components over disjoint parameters)
    if ((h->D <= 64 && h->fast_ready) || !stream_step_supported(h->M, h->likes.size())) return 0;
    if (!h->exts.empty()) return 0;  // external functions / priors: kernels_ext.cuh only
    const int D = h->D, NT = (D + 7) / 8, DP = 8 * NT;
    const int NL = (int)h->likes.size();
    // whitened coordinate a = aoff[l] + a' for component l; every sampled parameter must be
    // an input of exactly one component
    std::vector<int> like_of_i(D, -1), il_of_i(D, -1), aoff(NL + 1, 0);
    int nm = 1;
    for (int l = 0; l < NL; ++l) {
        const LikeHost &L = h->likes[l];
        aoff[l + 1] = aoff[l] + L.d.dim;
        nm = std::max(nm, (int)L.d.n_modes);
        for (int a = 0; a < L.d.dim; ++a) {
            if (like_of_i[L.idx[a]] != -1) return 0;  // shared / repeated parameter: general path
            like_of_i[L.idx[a]] = l;
            il_of_i[L.idx[a]] = a;
        }
    }
    for (int i = 0; i < D; ++i)
        if (like_of_i[i] < 0) return 0;
    StreamPackDesc P;
    memset(&P, 0, sizeof(P));
    P.NT = NT; P.DP = DP; P.n_modes = nm; P.n_like = NL;
    for (int l = 0; l < NL; ++l) P.like_modes[l] = h->likes[l].d.n_modes;
    // embedded matrices A^m[a][j] (block diagonal over the components); a component with
    // fewer modes repeats its last one (never read by the accept kernel)
    std::vector<std::vector<double>> Am(nm, std::vector<double>((size_t)DP * DP, 0.0));
    bool tri = true;
    for (int km = 0; km < nm; ++km)
        for (int j = 0; j < D; ++j) {
            const int i = h->i_of_j[j], l = like_of_i[i], il = il_of_i[i];
            const LikeHost &L = h->likes[l];
            const int d = L.d.dim, kl = std::min(km, (int)L.d.n_modes - 1);
            for (int a = 0; a < d; ++a) {
                const double v = L.linvT[(size_t)kl * d * d + (size_t)il * d + a];
                Am[km][(size_t)(aoff[l] + a) * DP + j] = v;
                if (aoff[l] + a < j && v != 0.0) tri = false;
            }
        }
    const bool frag = D <= CB2_STREAM_FRAG_MAX_D;  // above: plain matrices for k_rowgemm
    P.tri_like = tri ? 1 : 0;
    P.blocks_T = frag ? NT * (NT + 1) / 2 : 0;
    P.blocks_A = frag ? (tri ? P.blocks_T : NT * NT) : 0;
    int o = 0;
    auto take = [&](int n) { int r = o; o += (n + 1) & ~1; return r; };
    P.off_T = take(P.blocks_T * 64);
    P.off_A = take(nm * P.blocks_A * 64);
    P.off_mu = take(nm * DP);
    P.off_c0 = take(CB2_STREAM_MAX_LIKES * CB2_STREAM_MAX_MODES);
    P.off_w = take(CB2_STREAM_MAX_LIKES * CB2_STREAM_MAX_MODES);
    P.off_likeof = take(DP / 2 + 1);
    P.off_lower = take(DP); P.off_upper = take(DP); P.off_loc = take(DP);
    P.off_mls = take(DP); P.off_isc = take(DP); P.off_pa = take(DP); P.off_pb = take(DP);
    P.off_kind = take(DP);
    P.off_d1 = take(h->n_blocks * DP);
    P.off_w1 = take(h->n_blocks * nm * DP);
    P.total = o;
    {
        bool ident = (D == DP);
        for (int j = 0; j < D; ++j) ident = ident && (h->i_of_j[j] == j);
        P.iofj_identity = ident ? 1 : 0;
    }
    std::vector<double> pk(P.total, 0.0);
    std::vector<double> Tm((size_t)DP * DP, 0.0);
    for (int j = 0; j < D; ++j)
        for (int k = 0; k <= j; ++k) Tm[(size_t)j * DP + k] = h->Trow[(size_t)j * D + k];
    if (frag) pack_frag(pk, P.off_T, Tm, DP, NT, true);
    else {
        // row-major DP x DP copies for k_rowgemm: Trm[j][i] = T[j][i], Arm[a][j] = (L^-1 P)[a][j]
        int rcu;
        if ((rcu = upload(h, h->d_Trm, Tm))) return rcu;
        if ((rcu = upload(h, h->d_Arm, Am[0]))) return rcu;
    }
    for (int km = 0; km < nm; ++km) {
        if (frag) pack_frag(pk, P.off_A + (size_t)km * P.blocks_A * 64, Am[km], DP, NT, tri);
        for (int j = 0; j < D; ++j) {
            const int i = h->i_of_j[j], l = like_of_i[i];
            const LikeHost &L = h->likes[l];
            const int kl = std::min(km, (int)L.d.n_modes - 1);
            pk[P.off_mu + km * DP + j] = L.means[(size_t)kl * L.d.dim + il_of_i[i]];
        }
        // 1-parameter blocks: delta = T[:, j0] (RandProposer1D, proposal.py:86-93) and its image
        for (int b = 0; b < h->n_blocks; ++b) {
            if (h->bsize[b] != 1) continue;
            const int j0 = h->jstart[b];
            for (int j = 0; j < D; ++j) pk[P.off_d1 + (size_t)b * DP + j] = Tm[(size_t)j * DP + j0];
            for (int a = 0; a < D; ++a) {
                double acc = 0.0;
                for (int j = 0; j < D; ++j)
                    acc += Am[km][(size_t)a * DP + j] * Tm[(size_t)j * DP + j0];
                pk[P.off_w1 + ((size_t)b * nm + km) * DP + a] = acc;
            }
        }
    }
    int32_t *ilikeof = reinterpret_cast<int32_t *>(pk.data() + P.off_likeof);
    for (int a = 0; a < DP; ++a) ilikeof[a] = -1;
    for (int l = 0; l < NL; ++l) {
        const LikeHost &L = h->likes[l];
        for (int a = aoff[l]; a < aoff[l + 1]; ++a) ilikeof[a] = l;
        for (int km = 0; km < (int)L.d.n_modes; ++km) {
            pk[P.off_c0 + l * CB2_STREAM_MAX_MODES + km] = L.c0[km];
            pk[P.off_w + l * CB2_STREAM_MAX_MODES + km] = L.w[km];
        }
    }
    int32_t *ikind = reinterpret_cast<int32_t *>(pk.data() + P.off_kind);
    for (int j = 0; j < DP; ++j) {
        if (j < D) {
            const int i = h->i_of_j[j];
            pk[P.off_lower + j] = h->lower[i];
            pk[P.off_upper + j] = h->upper[i];
            pk[P.off_loc + j] = h->loc[i];
            pk[P.off_isc + j] = h->pscale[i];
            const int kd = h->prior_kind[i];
            pk[P.off_mls + j] = (kd == 1) ? (-std::log(h->pscale[i]) - CB2_LOG_2PI / 2)
                                : (kd >= 2 ? h->pcn[i] : 0.0);
            pk[P.off_pa + j] = kd >= 2 ? h->pa[i] : 0.0;
            pk[P.off_pb + j] = kd >= 2 ? h->pb[i] : 0.0;
            ikind[2 * j] = kd;
            ikind[2 * j + 1] = i;
        } else {
            pk[P.off_lower + j] = -INFINITY;
            pk[P.off_upper + j] = INFINITY;
            pk[P.off_isc + j] = 1.0;
            ikind[2 * j] = 0;
            ikind[2 * j + 1] = -1;
        }
    }
    int rc = upload(h, h->d_streampack, pk);
    if (rc) return rc;
    h->stream_desc = P;
    h->stream_ready = true;
    return 0;
}
