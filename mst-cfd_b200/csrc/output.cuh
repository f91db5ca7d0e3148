// output.cuh -- device half of the output path (included by mstgpu.cu).
// R = /root/reference/MST-CFD.
//
// Replaces the arithmetic of Work::writedataRhoBasedMshNodePlt (R/work/Work.cpp:243-304), which the
// reference runs on the host every 10 steps after reading the whole cell state back:
//   Work.cpp:248-285  face state: eta Q[c0] + (1 - eta) Q[c1] on interior faces, Q[c0] on boundary faces
//   Work.cpp:288-295  node state: sum_f w Qf / sum_f w over the node's faces in Node::addNbFace order,
//                     w = 1 / area(face[NODE id]) (the reference indexes the face list with the node id)
//   Work.cpp:299-303  rho, u_i, getT, getP, getMa (R/work/FUNCTION.cpp:8-20; 3-D adds w: extension)
// One thread per node; the face array of the reference is never formed; only [nnodes][D+4] doubles
// cross PCIe instead of the [ncells][D+2] state (tets: 56 B per node against 240 B of state per node).
//
// The arithmetic is written with the round-to-nearest intrinsics (__dmul_rn, __dadd_rn, __ddiv_rn,
// __dsqrt_rn), which nvcc never contracts into FMAs, in the reference's order of operations: the node
// fields are BIT-IDENTICAL to the reference's, so the 15-digit text file is identical byte for byte
// (tests/test_output_gpu.py).  Memory-bound and tiny next to a step (it runs once per 10 steps).
#pragma once

namespace {

template <int D>
__global__ void __launch_bounds__(256) k_node_fields(int nn, const int32_t* __restrict__ nf_ptr,
                                                     const int32_t* __restrict__ nf_idx, const int32_t* __restrict__ oc0,
                                                     const int32_t* __restrict__ oc1, const double* __restrict__ oeta,
                                                     const double* __restrict__ nw, const double* __restrict__ Q,
                                                     double gamma, double cv, double* __restrict__ out) {
    constexpr int U = D + 2, W = D + 4;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nn) return;
    double acc[U], d = 0.0;
#pragma unroll
    for (int k = 0; k < U; k++) acc[k] = 0.0;  // VCTDIMU::Zero(), Work.cpp:289
    const double w = nw[i];
    const int e = nf_ptr[i + 1];
    for (int j = nf_ptr[i]; j < e; j++) {
        const int f = nf_idx[j];
        const int a = oc0[f], b = oc1[f];
        double qf[U];
#pragma unroll
        for (int k = 0; k < U; k++) qf[k] = Q[(size_t)a * U + k];
        if (b >= 0) {  // Work.cpp:253-254
            const double et = oeta[f], om = __dsub_rn(1.0, et);
#pragma unroll
            for (int k = 0; k < U; k++) qf[k] = __dadd_rn(__dmul_rn(et, qf[k]), __dmul_rn(om, Q[(size_t)b * U + k]));
        }
#pragma unroll
        for (int k = 0; k < U; k++) acc[k] = __dadd_rn(acc[k], __dmul_rn(w, qf[k]));  // Work.cpp:292
        d = __dadd_rn(d, w);                                                           // Work.cpp:293
    }
    double q[U];
#pragma unroll
    for (int k = 0; k < U; k++) q[k] = __ddiv_rn(acc[k], d);  // Work.cpp:295
    double m2 = __dadd_rn(__dmul_rn(q[1], q[1]), __dmul_rn(q[2], q[2]));
    if (D == 3) m2 = __dadd_rn(m2, __dmul_rn(q[3], q[3]));
    const double ek = __ddiv_rn(__dmul_rn(0.5, m2), q[0]);          // 0.5 * (...) / rho
    const double ei = __dsub_rn(q[U - 1], ek);
    const double T = __ddiv_rn(__ddiv_rn(ei, q[0]), cv);            // FUNCTION.cpp:8-11
    const double p = __dmul_rn(ei, __dsub_rn(gamma, 1.0));          // FUNCTION.cpp:12-15
    const double Ma = __dsqrt_rn(__ddiv_rn(m2, __dmul_rn(__dmul_rn(gamma, p), q[0])));  // FUNCTION.cpp:16-20
    double* o = out + (size_t)i * W;
    o[0] = q[0];
#pragma unroll
    for (int k = 0; k < D; k++) o[1 + k] = __ddiv_rn(q[1 + k], q[0]);  // Work.cpp:301
    o[D + 1] = T;
    o[D + 2] = p;
    o[D + 3] = Ma;
}

void output_free(mstgpu_ctx* ctx) {
    for (void* q : {(void*)ctx->out_nf_ptr, (void*)ctx->out_nf_idx, (void*)ctx->out_c0, (void*)ctx->out_c1, (void*)ctx->out_eta,
                    (void*)ctx->out_w, (void*)ctx->out_fields, (void*)ctx->out_Q, (void*)ctx->out_send_idx, (void*)ctx->out_sendbuf})
        if (q) cudaFree(q);
    ctx->out_nf_ptr = ctx->out_nf_idx = ctx->out_c0 = ctx->out_c1 = ctx->out_send_idx = nullptr;
    ctx->out_eta = ctx->out_w = ctx->out_fields = ctx->out_Q = ctx->out_sendbuf = nullptr;
    ctx->out_nn = 0; ctx->out_nextra = 0; ctx->out_send_total = 0;
    ctx->out_node_ids.clear(); ctx->out_nbrs.clear();
}

}  // namespace

extern "C" {

int mstgpu_output_setup(mstgpu_ctx* ctx, const mstgpu_mesh* mesh, int32_t nnodes, const int32_t* nf_ptr,
                        const int32_t* nf_idx, const double* node_weight) {
    if (!ctx || !mesh || !nf_ptr || !nf_idx || nnodes <= 0) return MSTGPU_ERR_ARG;
    if (ctx->partitioned) { set_error(ctx, "output_setup: a partitioned context takes mstgpu_output_setup_partitioned"); return MSTGPU_ERR_STATE; }
    if (mesh->ncells != ctx->nc || mesh->nfaces != ctx->nf) { set_error(ctx, "output_setup: not the mesh this context was created from"); return MSTGPU_ERR_ARG; }
    CK(cudaSetDevice(ctx->device));
    const int nf = mesh->nfaces;
    const int64_t nnz = nf_ptr[nnodes];
    for (int64_t j = 0; j < nnz; j++)
        if (nf_idx[j] < 0 || nf_idx[j] >= nf) { set_error(ctx, "output_setup: face id out of range"); return MSTGPU_ERR_ARG; }
    // faces keep the reference's numbering here; cells go to the device order
    std::vector<int32_t> c0((size_t)nf), c1((size_t)nf);
    std::vector<double> eta((size_t)nf), w((size_t)nnodes);
    const auto& o2n = ctx->plan.cell_old2new;
#pragma omp parallel for schedule(static)
    for (int f = 0; f < nf; f++) {
        c0[(size_t)f] = o2n[(size_t)mesh->c0[f]];
        // only type-2 zones interpolate (Work.cpp:251-257); every other zone takes Q[c0].  Zone types the
        // reference's switch does not list (symmetry) leave its array uninitialised; Q[c0] here.
        const bool interp = mesh->ftype[f] == MSTGPU_BC_INTERIOR && mesh->c1[f] >= 0;
        c1[(size_t)f] = interp ? o2n[(size_t)mesh->c1[f]] : -1;
        eta[(size_t)f] = mesh->eta[f];
    }
    for (int i = 0; i < nnodes; i++) w[(size_t)i] = node_weight ? node_weight[i] : 1.0;
    // a second call replaces the tables (nothing dangles if an upload below fails)
    output_free(ctx);
    ctx->out_nn = nnodes;
    int r;
    if ((r = upload(ctx, &ctx->out_nf_ptr, std::vector<int32_t>(nf_ptr, nf_ptr + nnodes + 1)))) return r;
    if ((r = upload(ctx, &ctx->out_nf_idx, std::vector<int32_t>(nf_idx, nf_idx + nnz)))) return r;
    if ((r = upload(ctx, &ctx->out_c0, c0))) return r;
    if ((r = upload(ctx, &ctx->out_c1, c1))) return r;
    if ((r = upload(ctx, &ctx->out_eta, eta))) return r;
    if ((r = upload(ctx, &ctx->out_w, w))) return r;
    if ((r = dalloc(ctx, &ctx->out_fields, (size_t)nnodes * (ctx->D + 4)))) return r;
    CK(cudaStreamSynchronize(ctx->stream));
    return MSTGPU_OK;
}

// Output on a partitioned context.  The faces around a node on a partition cut belong to cells that can be several
// face hops apart, so the two face-ghost layers of the step do not hold them (and the ghost rows of the current buffer
// are one exchange old anyway).  Every node is computed by ONE rank -- the owner of c0 of the first face in its list --
// in the reference's face order, from fresh rows: the rows of other ranks' cells around its nodes are fetched when
// mstgpu_node_fields is called (one pack kernel + one grouped ncclSend / ncclRecv).  Every rank derives both its receive
// lists and its send lists from the global tables it already has (mesh, node lists, the cell -> part map kept in the
// partition), so no negotiation is needed; lists are sorted by global cell id on both sides.
int mstgpu_output_setup_partitioned(mstgpu_ctx* ctx, const mstgpu_part* part, const mstgpu_mesh* g, int32_t nnodes,
                                    const int32_t* nf_ptr, const int32_t* nf_idx, const double* node_weight) {
    if (!ctx || !part || !g || !nf_ptr || !nf_idx || nnodes <= 0) return MSTGPU_ERR_ARG;
    const Partition& P = part->p;
    if (!ctx->partitioned || ctx->n_owned != P.n_owned || ctx->nc != P.n_local) { set_error(ctx, "output_setup_partitioned: not the partition this context was created from"); return MSTGPU_ERR_ARG; }
    if ((int64_t)P.cell_part.size() != (int64_t)g->ncells) { set_error(ctx, "output_setup_partitioned: not the global mesh the partition was cut from"); return MSTGPU_ERR_ARG; }
    CK(cudaSetDevice(ctx->device));
    const int me = P.rank, np = P.nparts, ngf = g->nfaces;
    const int64_t nnz = nf_ptr[nnodes];
    for (int64_t j = 0; j < nnz; j++)
        if (nf_idx[j] < 0 || nf_idx[j] >= ngf) { set_error(ctx, "output_setup_partitioned: face id out of range"); return MSTGPU_ERR_ARG; }
    const int32_t* cp = P.cell_part.data();
    auto interp = [&](int f) { return g->ftype[f] == MSTGPU_BC_INTERIOR && g->c1[f] >= 0; };
    // cells of rank b around nodes of rank a (a != b): this rank keeps what it receives (a = me) and what it sends (b = me)
    std::vector<std::vector<int32_t>> recv(np), send(np);
    std::vector<int32_t> mine;
    for (int i = 0; i < nnodes; i++) {
        if (nf_ptr[i + 1] <= nf_ptr[i]) continue;  // a node without faces is nobody's
        const int owner = cp[g->c0[nf_idx[nf_ptr[i]]]];
        if (owner == me) mine.push_back(i);
        for (int j = nf_ptr[i]; j < nf_ptr[i + 1]; j++) {
            const int f = nf_idx[j];
            const int cells[2] = {g->c0[f], interp(f) ? g->c1[f] : -1};
            for (int c : cells) {
                if (c < 0) continue;
                const int pc = cp[c];
                if (pc == owner) continue;
                if (owner == me) recv[pc].push_back(c);
                else if (pc == me) send[owner].push_back(c);
            }
        }
    }
    for (auto* lists : {&recv, &send})
        for (auto& v : *lists) { std::sort(v.begin(), v.end()); v.erase(std::unique(v.begin(), v.end()), v.end()); }
    // global cell -> row of out_Q: owned cells keep their device-order row, received cells follow behind the local state
    std::vector<int32_t> g2l((size_t)g->ncells, -1);
    for (int i = 0; i < P.n_owned; i++) g2l[(size_t)P.local2global[i]] = ctx->plan.cell_old2new[i];
    output_free(ctx);
    int nextra = 0;
    std::vector<int32_t> sidx;
    for (int r = 0; r < np; r++) {
        if (recv[r].empty() && send[r].empty()) continue;
        mstgpu_ctx::OutNb nb{r, (int)sidx.size(), (int)send[r].size(), nextra, (int)recv[r].size()};
        for (int32_t c : send[r]) {
            if (g2l[(size_t)c] < 0 || cp[c] != me) { set_error(ctx, "output_setup_partitioned: internal: send list holds a cell this rank does not own"); return MSTGPU_ERR_ARG; }
            sidx.push_back(g2l[(size_t)c]);
        }
        for (int32_t c : recv[r]) g2l[(size_t)c] = ctx->nc + nextra++;
        ctx->out_nbrs.push_back(nb);
    }
    // compact face list of my nodes (global face ids -> 0..), cells as rows of out_Q
    std::vector<int32_t> fmap((size_t)ngf, -1), c0, c1, lptr((size_t)mine.size() + 1, 0), lidx;
    std::vector<double> eta, w(mine.size());
    for (size_t k = 0; k < mine.size(); k++) {
        const int i = mine[k];
        for (int j = nf_ptr[i]; j < nf_ptr[i + 1]; j++) {
            const int f = nf_idx[j];
            if (fmap[(size_t)f] < 0) {
                fmap[(size_t)f] = (int32_t)c0.size();
                const int a = g2l[(size_t)g->c0[f]], b = interp(f) ? g2l[(size_t)g->c1[f]] : -1;
                if (a < 0 || (interp(f) && b < 0)) { set_error(ctx, "output_setup_partitioned: internal: a cell around an owned node has no row"); return MSTGPU_ERR_ARG; }
                c0.push_back(a); c1.push_back(b); eta.push_back(g->eta[f]);
            }
            lidx.push_back(fmap[(size_t)f]);
        }
        lptr[k + 1] = (int32_t)lidx.size();
        w[k] = node_weight ? node_weight[i] : 1.0;
    }
    ctx->out_nn = (int)mine.size();
    ctx->out_node_ids = mine;
    ctx->out_nextra = nextra;
    ctx->out_send_total = (int)sidx.size();
    int r;
    if ((r = upload(ctx, &ctx->out_nf_ptr, lptr))) return r;
    if ((r = upload(ctx, &ctx->out_nf_idx, lidx))) return r;
    if ((r = upload(ctx, &ctx->out_c0, c0))) return r;
    if ((r = upload(ctx, &ctx->out_c1, c1))) return r;
    if ((r = upload(ctx, &ctx->out_eta, eta))) return r;
    if ((r = upload(ctx, &ctx->out_w, w))) return r;
    if ((r = upload(ctx, &ctx->out_send_idx, sidx))) return r;
    if ((r = dalloc(ctx, &ctx->out_fields, std::max<size_t>(1, (size_t)ctx->out_nn * (ctx->D + 4))))) return r;
    if ((r = dalloc(ctx, &ctx->out_Q, (size_t)(ctx->nc + nextra) * ctx->U))) return r;
    if ((r = dalloc(ctx, &ctx->out_sendbuf, std::max<size_t>(1, sidx.size() * ctx->U)))) return r;
    CK(cudaStreamSynchronize(ctx->stream));
    return MSTGPU_OK;
}

int32_t mstgpu_output_node_count(mstgpu_ctx* ctx) { return ctx ? ctx->out_nn : -1; }

int mstgpu_output_node_ids(mstgpu_ctx* ctx, int32_t* ids) {
    if (!ctx || !ids) return MSTGPU_ERR_ARG;
    if (ctx->out_node_ids.empty()) for (int i = 0; i < ctx->out_nn; i++) ids[i] = i;
    else std::memcpy(ids, ctx->out_node_ids.data(), sizeof(int32_t) * ctx->out_node_ids.size());
    return MSTGPU_OK;
}

int mstgpu_node_fields(mstgpu_ctx* ctx, double* out) {
    if (!ctx || !out) return MSTGPU_ERR_ARG;
    if (!ctx->out_fields) { set_error(ctx, "node_fields before output_setup"); return MSTGPU_ERR_STATE; }
    if (!ctx->has_state) { set_error(ctx, "node_fields before set_state"); return MSTGPU_ERR_STATE; }
    CK(cudaSetDevice(ctx->device));
    const int nn = ctx->out_nn;
    const unsigned grid = (unsigned)((std::max(nn, 1) + 255) / 256);
    const double* Qsrc = ctx->Q[ctx->cur];
    if (ctx->out_Q) {
        // partitioned: the local state + fresh rows of the other ranks' cells around this rank's nodes
        const int U = ctx->U;
        if (!ctx->out_nbrs.empty() && !ctx->comm) { set_error(ctx, "partitioned context without a communicator: call mstgpu_comm_init"); return MSTGPU_ERR_STATE; }
        CK(cudaMemcpyAsync(ctx->out_Q, Qsrc, (size_t)ctx->nc * U * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
        if (ctx->out_send_total > 0) {
            k_pack_rows<<<(ctx->out_send_total * U + 255) / 256, 256, 0, ctx->stream>>>(ctx->out_send_total, U, ctx->out_send_idx, Qsrc, ctx->out_sendbuf);
            ctx->launches++;
        }
        if (!ctx->out_nbrs.empty()) {
            NK(g_nccl.GroupStart());
            for (const auto& nb : ctx->out_nbrs) {
                if (nb.send_count) NK(g_nccl.Send(ctx->out_sendbuf + (size_t)nb.send_off * U, (size_t)nb.send_count * U, ncclDouble, nb.rank, ctx->comm, ctx->stream));
                if (nb.recv_count) NK(g_nccl.Recv(ctx->out_Q + (size_t)(ctx->nc + nb.recv_off) * U, (size_t)nb.recv_count * U, ncclDouble, nb.rank, ctx->comm, ctx->stream));
            }
            NK(g_nccl.GroupEnd());
        }
        Qsrc = ctx->out_Q;
    }
    if (nn > 0) {
        KTimer t(ctx, "node_fields");
        if (ctx->D == 2)
            k_node_fields<2><<<grid, 256, 0, ctx->stream>>>(nn, ctx->out_nf_ptr, ctx->out_nf_idx, ctx->out_c0, ctx->out_c1, ctx->out_eta,
                                                            ctx->out_w, Qsrc, ctx->cfg.gamma, ctx->cfg.cv, ctx->out_fields);
        else
            k_node_fields<3><<<grid, 256, 0, ctx->stream>>>(nn, ctx->out_nf_ptr, ctx->out_nf_idx, ctx->out_c0, ctx->out_c1, ctx->out_eta,
                                                            ctx->out_w, Qsrc, ctx->cfg.gamma, ctx->cfg.cv, ctx->out_fields);
    }
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, ctx->out_fields, (size_t)nn * (ctx->D + 4) * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (ctx->ktiming) drain_timers(ctx);
    return MSTGPU_OK;
}

}  // extern "C"
