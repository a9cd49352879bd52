// K2: level-synchronous hydration of the hash-consed store DAG.
//
// Replaces StoreCore::hydrate_z_cache / hash_ptr_val_unsafe (reference src/lem/store_core.rs:199-269), which walks
// the DAG recursively on the CPU (rayon only inside chunks of 256), with: (host) one pass over the topologically
// ordered node list to assign each node its height, a counting sort by (height, arity); (device) per height and
// arity one gather of the children's digests into flat preimages -- layouts of `impl StoreHasher for PoseidonCache`,
// src/lem/store.rs:29-78 -- one batch Poseidon launch, one scatter of the digests into the digest table.  All
// nodes of one height are independent.  Digest table and preimages stay in Montgomery form on the device.
#include "poseidon_api.h"

#include <algorithm>

namespace lurk {

// node -> preimage row; tag -> field element is F::from(u16) (src/tag.rs:99-101)
template <class F>
__global__ void __launch_bounds__(128) dag_gather_kernel(const lurk_dag_node *__restrict__ nodes, const uint32_t *__restrict__ batch,
                                                         uint32_t count, int arity, const F *__restrict__ table, F *__restrict__ pre) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const lurk_dag_node nd = nodes[batch[i]];
    F *o = pre + (size_t)i * arity;
    if (nd.kind <= LURK_DAG_TUPLE4) {
        const int nch = nd.kind;
        for (int c = 0; c < nch; c++) {
            store_fe(o + 2 * c, F::from_u64(nd.tag[c]));
            store_fe(o + 2 * c + 1, load_fe<F>(table + nd.child[c]));
        }
    } else if (nd.kind == LURK_DAG_COMPACT) {
        store_fe(o + 0, load_fe<F>(table + nd.child[0]));
        store_fe(o + 1, F::from_u64(nd.tag[1]));
        store_fe(o + 2, load_fe<F>(table + nd.child[1]));
        store_fe(o + 3, load_fe<F>(table + nd.child[2]));
    } else {   // commitment
        store_fe(o + 0, load_fe<F>(table + nd.child[0]));
        store_fe(o + 1, F::from_u64(nd.tag[1]));
        store_fe(o + 2, load_fe<F>(table + nd.child[1]));
    }
}
template <class F>
__global__ void __launch_bounds__(128) dag_scatter_kernel(const F *__restrict__ digests, const uint32_t *__restrict__ batch, uint32_t count,
                                                          uint32_t n_atoms, F *__restrict__ table) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    store_fe(table + n_atoms + batch[i], load_fe<F>(digests + i));
}

static int kind_arity(int kind) {
    switch (kind) {
        case LURK_DAG_TUPLE2: return 4;
        case LURK_DAG_TUPLE3: return 6;
        case LURK_DAG_TUPLE4: return 8;
        case LURK_DAG_COMPACT: return 4;
        case LURK_DAG_COMMITMENT: return 3;
    }
    return 0;
}
static int kind_children(int kind) {
    switch (kind) {
        case LURK_DAG_TUPLE2: return 2;
        case LURK_DAG_TUPLE3: return 3;
        case LURK_DAG_TUPLE4: return 4;
        case LURK_DAG_COMPACT: return 3;
        case LURK_DAG_COMMITMENT: return 2;
    }
    return 0;
}

template <class F>
static int dag_hash(const lurk_dag_node *nodes, size_t n, const uint8_t *atoms, size_t n_atoms, uint8_t *out) {
    if (n + n_atoms >= 0xffffffffull) { set_error("DAG too large"); return LURK_ERR_ARG; }
    // ---- host: heights and (height, arity) batches
    std::vector<uint32_t> height(n);
    uint32_t max_h = 0;
    for (size_t i = 0; i < n; i++) {
        const int nch = kind_children(nodes[i].kind);
        if (!nch) { set_error("node %zu: unknown kind %d", i, (int)nodes[i].kind); return LURK_ERR_ARG; }
        uint32_t h = 0;
        for (int c = 0; c < nch; c++) {
            uint64_t ix = nodes[i].child[c];
            if (ix < n_atoms) continue;
            ix -= n_atoms;
            if (ix >= i) { set_error("node %zu refers to node %llu which does not precede it", i, (unsigned long long)ix); return LURK_ERR_ORDER; }
            h = std::max(h, height[ix] + 1);
        }
        height[i] = h;
        max_h = std::max(max_h, h);
    }
    static const int ARITIES[4] = {3, 4, 6, 8};
    auto slot = [&](size_t i) { int a = kind_arity(nodes[i].kind); return (size_t)height[i] * 4 + (a == 3 ? 0 : a == 4 ? 1 : a == 6 ? 2 : 3); };
    std::vector<uint32_t> start(((size_t)max_h + 1) * 4 + 1, 0);
    for (size_t i = 0; i < n; i++) start[slot(i) + 1]++;
    for (size_t k = 1; k < start.size(); k++) start[k] += start[k - 1];
    std::vector<uint32_t> order(n), fill(start.begin(), start.end() - 1);
    size_t max_batch = 0;
    for (size_t i = 0; i < n; i++) order[fill[slot(i)]++] = (uint32_t)i;
    for (size_t k = 0; k + 1 < start.size(); k++) max_batch = std::max<size_t>(max_batch, start[k + 1] - start[k]);

    // ---- device
    DevBuf d_nodes, d_order, d_table, d_pre, d_dig;
    LURK_TRY(d_nodes.alloc(n * sizeof(lurk_dag_node)));
    LURK_TRY(d_order.alloc(n * sizeof(uint32_t)));
    LURK_TRY(d_table.alloc((n_atoms + n) * sizeof(F)));
    LURK_TRY(d_pre.alloc(max_batch * 8 * sizeof(F)));
    LURK_TRY(d_dig.alloc(max_batch * sizeof(F)));
    LURK_CUDA_TRY(cudaMemcpy(d_nodes.p, nodes, d_nodes.bytes, cudaMemcpyHostToDevice));
    LURK_CUDA_TRY(cudaMemcpy(d_order.p, order.data(), d_order.bytes, cudaMemcpyHostToDevice));
    if (n_atoms) {
        LURK_CUDA_TRY(cudaMemcpy(d_table.p, atoms, n_atoms * sizeof(F), cudaMemcpyHostToDevice));
        int bad = 0;
        LURK_TRY(check_reduced_dev<F>(d_table.p, n_atoms, 0, &bad));
        if (bad) { set_error("%d atom digest(s) are not reduced below the field modulus", bad); return LURK_ERR_RANGE; }
        LURK_TRY(convert_dev<F>(d_table.p, n_atoms, LURK_FMT_MONTGOMERY, d_table.p, 0));
    }
    for (uint32_t h = 0; h <= max_h; h++) {
        for (int a = 0; a < 4; a++) {
            const uint32_t b0 = start[(size_t)h * 4 + a], cnt = start[(size_t)h * 4 + a + 1] - b0;
            if (!cnt) continue;
            const uint32_t *batch = d_order.as<uint32_t>() + b0;
            dag_gather_kernel<F><<<(cnt + 127) / 128, 128>>>(d_nodes.as<lurk_dag_node>(), batch, cnt, ARITIES[a], d_table.as<F>(), d_pre.as<F>());
            LURK_TRY((launch_poseidon<F, false>(ARITIES[a], d_pre.p, cnt, d_dig.p, LURK_FMT_MONTGOMERY, LURK_FMT_MONTGOMERY, 0)));
            dag_scatter_kernel<F><<<(cnt + 127) / 128, 128>>>(d_dig.as<F>(), batch, cnt, (uint32_t)n_atoms, d_table.as<F>());
        }
    }
    LURK_CUDA_TRY(cudaGetLastError());
    F *node_table = d_table.as<F>() + n_atoms;
    LURK_TRY(convert_dev<F>(node_table, n, LURK_FMT_CANONICAL, node_table, 0));
    LURK_CUDA_TRY(cudaMemcpy(out, node_table, n * sizeof(F), cudaMemcpyDeviceToHost));
    return LURK_OK;
}

}  // namespace lurk

using namespace lurk;

extern "C" int lurk_dag_hash(int field_id, const lurk_dag_node *nodes, size_t n, const uint8_t *atom_digests, size_t n_atoms,
                             uint8_t *out_digests) {
    LURK_TRY(require_gpu());
    if (n == 0) return LURK_OK;
    if (!nodes || !out_digests || (n_atoms && !atom_digests)) { set_error("null argument"); return LURK_ERR_ARG; }
    return dispatch_field(field_id, [&](auto f) { return dag_hash<decltype(f)>(nodes, n, atom_digests, n_atoms, out_digests); });
}
