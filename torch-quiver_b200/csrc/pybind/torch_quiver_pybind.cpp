// torch_quiver_pybind.cpp -- the reference's pybind11 plugin surface (module `torch_quiver`,
// srcs/cpp/src/quiver/torch/module.cpp:16-26; sampler bindings srcs/cpp/src/quiver/cuda/quiver_sample.cu:500-513; feature
// bindings srcs/cpp/src/quiver/cuda/quiver_feature.cu:431-473) implemented as a thin adapter over the C ABI of
// libquiver_b200.so (include/quiver_b200.h).  Same class and method names, argument order and return shapes as the
// reference, so its Python package (srcs/python/quiver) can import this module in place of its own extension.
//
// Built by torch-quiver_b200/csrc/pybind/build.py into torch-quiver_b200/torch_quiver_pybind/torch_quiver*.so (g++ against
// this image's torch headers; links -lquiver_b200).  The ctypes adapter torch-quiver_b200/torch_quiver/__init__.py is the
// same mapping without a torch-linked C++ build; tests/test_gpu_pybind_adapter.py runs parity checks through THIS one.
//
// Differences from the reference that a caller can observe are listed in INTEGRATION.md ("Behavioural differences"):
// errors raise instead of exit(1), work runs on torch's current stream, invalid ids give zero rows, shards are freed.
#include <torch/extension.h>

#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>

#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "quiver_b200.h"

namespace
{
void ok(int rc)
{
    if (rc != QV_OK) throw std::runtime_error(std::string("libquiver_b200 error ") + std::to_string(rc) + ": " + qv_last_error());
}

void *cur_stream(int device) { return at::cuda::getCurrentCUDAStream(device).stream(); }

const int64_t *long_ptr(const torch::Tensor &t, const char *name, int device)
{
    if (t.scalar_type() != torch::kInt64) throw std::runtime_error(std::string(name) + " must be a torch.long tensor");
    if (!t.is_cuda()) throw std::runtime_error(std::string(name) + " must be a CUDA tensor");
    if (device >= 0 && t.get_device() != device)
        throw std::runtime_error(std::string(name) + " lives on another device than this object");
    if (!t.is_contiguous()) throw std::runtime_error(std::string(name) + " must be contiguous");
    return t.data_ptr<int64_t>();
}

// ---- class TorchQuiver (quiver_sample.cu:77-357) ---------------------------------------------------------------------
class Quiver
{
  public:
    Quiver(torch::Tensor indptr, torch::Tensor indices, torch::Tensor edge_ids, int device, bool cuda) : device_(device)
    {
        TORCH_CHECK(indptr.dim() == 1 && indices.dim() == 1, "check_eq failed");  // quiver_sample.cu:373-376
        TORCH_CHECK(indptr.scalar_type() == torch::kInt64 && indices.scalar_type() == torch::kInt64,
                    "indptr / indices must be torch.long");
        TORCH_CHECK(indptr.numel() >= 1, "indptr must hold at least one entry");
        const auto dev = torch::Device(torch::kCUDA, device);
        indptr_ = indptr.to(dev).contiguous();  // always in HBM (quiver_sample.cu:401-407)
        const int64_t *idx = nullptr;
        if (cuda || indices.is_cuda()) {
            indices_ = indices.to(dev).contiguous();
            idx = indices_.data_ptr<int64_t>();
        } else {  // UVA / zero-copy: alias the caller's CPU tensor (quiver_sample.cu:413-421)
            indices_ = indices.contiguous();
            if (indices_.numel() > 0) {
                void *alias = nullptr;
                ok(qv_host_register(device, indices_.data_ptr(), indices_.numel() * 8, &alias));
                registered_.push_back(indices_.data_ptr());
                idx = static_cast<const int64_t *>(alias);
            }
        }
        ok(qv_sampler_create(device, indptr_.data_ptr<int64_t>(), indptr_.numel() - 1, idx, indices_.numel(), &h_));
        if (edge_ids.defined() && edge_ids.dim() == 1 && edge_ids.numel() == indices_.numel() && indices_.numel() > 0 &&
            edge_ids.scalar_type() == torch::kInt64) {  // use_eid, quiver_sample.cu:385-387
            const int64_t *eid = nullptr;
            if (cuda || edge_ids.is_cuda()) {
                edge_ids_ = edge_ids.to(dev).contiguous();
                eid = edge_ids_.data_ptr<int64_t>();
            } else {
                edge_ids_ = edge_ids.contiguous();
                void *alias = nullptr;
                ok(qv_host_register(device, edge_ids_.data_ptr(), edge_ids_.numel() * 8, &alias));
                registered_.push_back(edge_ids_.data_ptr());
                eid = static_cast<const int64_t *>(alias);
            }
            ok(qv_sampler_set_edge_ids(h_, eid));
        }
    }
    Quiver(const Quiver &) = delete;
    Quiver &operator=(const Quiver &) = delete;
    ~Quiver()
    {
        if (h_) qv_sampler_destroy(h_);
        for (void *p : registered_) qv_host_unregister(p);
    }

    // Quiver.sample_neighbor(stream_num, vertices, k) -> (neighbors, counts)   quiver_sample.cu:113
    std::tuple<torch::Tensor, torch::Tensor> sample_neighbor(int /*stream_num*/, const torch::Tensor &vertices, int64_t k)
    {
        const int64_t *v = long_ptr(vertices, "vertices", device_);
        c10::cuda::CUDAGuard guard(device_);
        const int64_t S = vertices.numel();
        auto counts = torch::empty({S}, vertices.options()), out_ptr = torch::empty({S}, vertices.options());
        int64_t total = 0;
        ok(qv_sample_count(h_, v, S, k, counts.data_ptr<int64_t>(), out_ptr.data_ptr<int64_t>(), &total, cur_stream(device_)));
        auto neighbors = torch::empty({total}, vertices.options());
        ok(qv_sample_fill(h_, v, S, k, rand_seed, out_ptr.data_ptr<int64_t>(), neighbors.data_ptr<int64_t>(), nullptr,
                          cur_stream(device_)));
        return {neighbors, counts};
    }

    // Quiver.reindex_single(inputs, outputs, counts) -> (frontier, row_idx, col_idx)   quiver_sample.cu:305
    std::tuple<torch::Tensor, torch::Tensor, torch::Tensor> reindex_single(torch::Tensor inputs, torch::Tensor outputs,
                                                                            torch::Tensor counts)
    {
        const int64_t *in = long_ptr(inputs, "inputs", device_), *out = long_ptr(outputs, "outputs", device_);
        const int64_t *cnt = long_ptr(counts, "counts", device_);
        TORCH_CHECK(counts.numel() == inputs.numel(), "counts must have one entry per input");
        c10::cuda::CUDAGuard guard(device_);
        const int64_t S = inputs.numel(), tot = outputs.numel();
        auto frontier = torch::empty({S + tot}, inputs.options());
        auto row = torch::empty({tot}, inputs.options()), col = torch::empty({tot}, inputs.options());
        int64_t F = 0;
        ok(qv_reindex(h_, in, S, out, tot, cnt, frontier.data_ptr<int64_t>(), row.data_ptr<int64_t>(), col.data_ptr<int64_t>(),
                      &F, cur_stream(device_)));
        return {frontier.narrow(0, 0, F), row, col};
    }

    // Quiver.sample_sub(stream_num, vertices, k) -> (frontier, row_idx, col_idx)   quiver_sample.cu:257 -- one fused call
    std::tuple<torch::Tensor, torch::Tensor, torch::Tensor> sample_sub(int stream_num, const torch::Tensor &vertices, int64_t k)
    {
        if (k >= 0 && vertices.numel() > 0) {
            int64_t bn[2], be[1];
            if (qv_khop_bounds(vertices.numel(), &k, 1, bn, be) == QV_OK) {
                const int64_t *v = long_ptr(vertices, "vertices", device_);
                c10::cuda::CUDAGuard guard(device_);
                auto frontier = torch::empty({std::max<int64_t>(bn[1], 1)}, vertices.options());
                auto edges = torch::empty({std::max<int64_t>(2 * be[0], 2)}, vertices.options());
                int64_t *ebuf = edges.data_ptr<int64_t>(), nodes[2] = {0, 0}, E[1] = {0};
                ok(qv_khop(h_, v, vertices.numel(), &k, 1, rand_seed, frontier.data_ptr<int64_t>(), &ebuf, nullptr, nodes, E,
                           cur_stream(device_)));
                return {frontier.narrow(0, 0, nodes[1]), edges.narrow(0, E[0], E[0]), edges.narrow(0, 0, E[0])};
            }
        }
        auto [out, cnt] = sample_neighbor(stream_num, vertices, k);
        return reindex_single(vertices, out, cnt);
    }

    // Quiver.cal_neighbor_prob(stream_num, last_prob, cur_prob, k)   quiver_sample.cu:100
    void cal_neighbor_prob(int /*stream_num*/, torch::Tensor last_prob, torch::Tensor cur_prob, int k)
    {
        TORCH_CHECK(last_prob.is_cuda() && cur_prob.is_cuda() && last_prob.scalar_type() == torch::kFloat32 &&
                        cur_prob.scalar_type() == torch::kFloat32 && last_prob.is_contiguous() && cur_prob.is_contiguous(),
                    "cal_neighbor_prob expects contiguous float32 CUDA tensors");
        c10::cuda::CUDAGuard guard(device_);
        ok(qv_cal_neighbor_prob(h_, last_prob.data_ptr<float>(), cur_prob.data_ptr<float>(), cur_prob.numel(), k,
                                cur_stream(device_)));
    }

    // Extension (ours): all hops of GraphSageSampler.sample (sage_sampler.py:118-147) in one C call.
    // Returns (n_id, [edge_index[2, E_l]], [[n_src_l, n_dst_l]]) innermost hop first.
    std::tuple<torch::Tensor, std::vector<torch::Tensor>, std::vector<std::vector<int64_t>>>
    sample_khop(const torch::Tensor &seeds, std::vector<int64_t> sizes)
    {
        const int n_hops = static_cast<int>(sizes.size());
        TORCH_CHECK(n_hops >= 1 && n_hops <= QV_MAX_HOPS, "between 1 and ", QV_MAX_HOPS, " hops");
        const int64_t *v = long_ptr(seeds, "seeds", device_);
        int64_t bn[QV_MAX_HOPS + 1], be[QV_MAX_HOPS];
        ok(qv_khop_bounds(seeds.numel(), sizes.data(), n_hops, bn, be));
        c10::cuda::CUDAGuard guard(device_);
        int64_t total = (std::max<int64_t>(bn[n_hops], 1) + 1) / 2 * 2, offs[QV_MAX_HOPS];
        for (int h = 0; h < n_hops; h++) {
            offs[h] = total;
            total += std::max<int64_t>(2 * be[h], 2);
        }
        auto arena = torch::empty({total}, seeds.options());
        int64_t *base = arena.data_ptr<int64_t>(), *bufs[QV_MAX_HOPS], nodes[QV_MAX_HOPS + 1], edges[QV_MAX_HOPS];
        for (int h = 0; h < n_hops; h++) bufs[h] = base + offs[h];
        ok(qv_khop(h_, v, seeds.numel(), sizes.data(), n_hops, rand_seed, base, bufs, nullptr, nodes, edges, cur_stream(device_)));
        std::vector<torch::Tensor> edge_index;
        std::vector<std::vector<int64_t>> hop_sizes;
        for (int h = 0; h < n_hops; h++) {
            edge_index.push_back(arena.narrow(0, offs[h], 2 * edges[h]).view({2, edges[h]}));
            hop_sizes.push_back({nodes[h + 1], nodes[h]});
        }
        return {arena.narrow(0, 0, nodes[n_hops]), edge_index, hop_sizes};
    }

    uint64_t rand_seed = 0;  // the reference hard-codes 0 (quiver.cu.hpp:392)

  private:
    int device_;
    qv_sampler *h_ = nullptr;
    torch::Tensor indptr_, indices_, edge_ids_;
    std::vector<void *> registered_;
};

std::shared_ptr<Quiver> device_quiver_from_csr_array(torch::Tensor indptr, torch::Tensor indices, torch::Tensor edge_ids,
                                                     int device, bool cuda)
{
    return std::make_shared<Quiver>(indptr, indices, edge_ids, device, cuda);
}

// ---- class ShardTensorItem (quiver_feature.cu:20-55) -------------------------------------------------------------------
struct ShardTensorItem {
    int device = -1;
    int element_size = 4;
    std::string mem_handle = std::string(QV_IPC_HANDLE_BYTES, '\0');
    std::vector<int64_t> shape;
    std::tuple<int, int, py::bytes, std::vector<int64_t>> share_ipc() { return {device, element_size, py::bytes(mem_handle), shape}; }
    void from_ipc(std::tuple<int, int, py::bytes, std::vector<int64_t>> t)
    {
        device = std::get<0>(t);
        element_size = std::get<1>(t);
        mem_handle = std::string(std::get<2>(t));
        shape = std::get<3>(t);
        TORCH_CHECK(mem_handle.size() == QV_IPC_HANDLE_BYTES, "a CUDA IPC handle has ", QV_IPC_HANDLE_BYTES, " bytes");
    }
};

// ---- class ShardTensor (quiver_feature.cu:57-376) ------------------------------------------------------------------------
class ShardTensor
{
    struct Shard {
        int device;       // owner GPU, -1 = pinned host
        void *ptr;        // valid on device_
        int64_t rows, pitch;
        bool owned, ipc_opened;
        void *host_base;  // what we registered (host tier)
        torch::Tensor keep;
        std::vector<int64_t> shape;
    };

  public:
    explicit ShardTensor(int device) : device_(device) {}
    ShardTensor(const ShardTensor &) = delete;
    ~ShardTensor()
    {
        for (auto &s : shards_) {
            if (s.owned && s.ptr)
                qv_free(s.device, s.ptr);
            else if (s.ipc_opened && s.ptr)
                qv_ipc_close_handle(device_, s.ptr);
            else if (s.host_base)
                qv_host_unregister(s.host_base);
        }
    }

    void append(torch::Tensor &tensor, int target_device)  // quiver_feature.cu:145-206
    {
        TORCH_CHECK(!tensor.is_cuda(), "tensor must be CPU tensor");  // CHECK_CPU, quiver_feature.cu:19,147
        auto x = tensor.contiguous();
        admit(x.sizes().vec(), static_cast<int>(x.element_size()));
        const int64_t rows = x.size(0), rb = row_bytes();
        Shard s{target_device, nullptr, rows, rb, false, false, nullptr, {}, x.sizes().vec()};
        if (target_device >= 0) {
            s.pitch = (rb + 15) / 16 * 16;
            ok(qv_malloc(target_device, static_cast<size_t>(std::max<int64_t>(rows * s.pitch, 16)), &s.ptr));
            s.owned = true;
            ok(qv_upload_rows(target_device, s.ptr, s.pitch, x.data_ptr(), rb, rb, rows));
            int can = 0;
            if (target_device != device_ && qv_can_device_access_peer(device_, target_device, &can) == QV_OK && can) {
                const int pair[2] = {device_, target_device};
                ok(qv_init_p2p(pair, 2, nullptr));
            }
        } else {  // zero-copy host tier: aliased, not copied (quiver_feature.cu:192-199) -- kept alive here
            s.keep = x;
            if (x.numel() > 0) {
                ok(qv_host_register(device_, x.data_ptr(), x.numel() * x.element_size(), &s.ptr));
                s.host_base = x.data_ptr();
            }
        }
        shards_.push_back(std::move(s));
    }

    void append_item(ShardTensorItem item)  // quiver_feature.cu:86-143
    {
        admit(item.shape, item.element_size);
        Shard s{item.device, nullptr, item.shape[0], (row_bytes() + 15) / 16 * 16, false, true, nullptr, {}, item.shape};
        ok(qv_ipc_open_handle(device_, reinterpret_cast<const unsigned char *>(item.mem_handle.data()), &s.ptr));
        shards_.push_back(std::move(s));
    }

    torch::Tensor getitem(torch::Tensor &indices)  // quiver_feature.cu:246-302
    {
        const int current = indices.is_cuda() ? indices.get_device() : device_;
        const int64_t *idx = long_ptr(indices, "indices", -1);
        TORCH_CHECK(!shards_.empty(), "ShardTensor is empty");
        qv_shard_table t{};
        t.n_shards = static_cast<int32_t>(shards_.size());
        for (size_t s = 0; s < shards_.size(); s++) {
            t.row_begin[s] = offsets_[s];
            t.ptr[s] = shards_[s].ptr;
            t.pitch[s] = shards_[s].pitch;
            int can = 1;
            if (shards_[s].device >= 0 && shards_[s].device != current) ok(qv_can_device_access_peer(current, shards_[s].device, &can));
            t.accessible[s] = can;
        }
        t.row_begin[shards_.size()] = offsets_.back();
        c10::cuda::CUDAGuard guard(current);
        std::vector<int64_t> out_shape(shape_);
        out_shape[0] = indices.numel();
        const auto dtype = element_size_ == 2 ? torch::kFloat16 : (element_size_ == 8 ? torch::kFloat64 : (element_size_ == 1 ? torch::kUInt8 : torch::kFloat32));
        auto res = torch::empty(out_shape, torch::TensorOptions().dtype(dtype).device(torch::kCUDA, current));
        ok(qv_gather(&t, idx, nullptr, indices.numel(), row_bytes(), res.data_ptr(), 0, cur_stream(current)));
        return res;
    }

    std::vector<ShardTensorItem> share_ipc()  // quiver_feature.cu:335-350
    {
        std::vector<ShardTensorItem> items;
        for (auto &s : shards_) {
            if (s.device < 0) continue;
            TORCH_CHECK(s.owned, "only the process that created a GPU shard can export it");
            ShardTensorItem it;
            it.device = s.device;
            it.element_size = element_size_;
            it.shape = s.shape;
            unsigned char h[QV_IPC_HANDLE_BYTES];
            ok(qv_ipc_get_handle(s.device, s.ptr, h));
            it.mem_handle.assign(reinterpret_cast<const char *>(h), QV_IPC_HANDLE_BYTES);
            items.push_back(std::move(it));
        }
        return items;
    }

    void unregister(torch::Tensor &cpu_tensor)  // quiver_feature.cu:354-360
    {
        ok(qv_host_unregister(cpu_tensor.data_ptr()));
        for (auto &s : shards_)
            if (s.host_base == cpu_tensor.data_ptr()) s.host_base = nullptr;
    }

    std::vector<int64_t> shape() const { return shape_; }
    int device() const { return device_; }
    int64_t size(int dim) const { return shape_.empty() ? 0 : shape_[dim]; }
    int64_t stride(int dim) const
    {
        int64_t r = 1;
        for (size_t d = dim + 1; d < shape_.size(); d++) r *= shape_[d];
        return r;
    }
    int64_t numel() const
    {
        int64_t r = 1;
        for (auto d : shape_) r *= d;
        return r;
    }
    int device_count() const { return static_cast<int>(shards_.size()); }

  private:
    int64_t row_bytes() const { return stride(0) * element_size_; }
    void admit(const std::vector<int64_t> &shape, int element_size)
    {
        TORCH_CHECK(!shape.empty(), "a shard needs at least one dimension");
        if (shape_.empty()) {
            shape_ = shape;
            shape_[0] = 0;
            element_size_ = element_size;
        } else {
            TORCH_CHECK(std::vector<int64_t>(shape.begin() + 1, shape.end()) == std::vector<int64_t>(shape_.begin() + 1, shape_.end()) &&
                            element_size == element_size_,
                        "shard shape / element size does not match the table");
        }
        TORCH_CHECK(shards_.size() < QV_MAX_SHARDS, "at most ", QV_MAX_SHARDS, " shards per ShardTensor");
        shape_[0] += shape[0];
        offsets_.push_back(offsets_.back() + shape[0]);
    }

    int device_;
    int element_size_ = 4;
    std::vector<int64_t> shape_;
    std::vector<int64_t> offsets_{0};
    std::vector<Shard> shards_;
};

// ---- compiled call path of the ctypes package's fused k-hop (ours) -------------------------------------------------------
// torch_quiver.Quiver.sample_khop (torch_quiver/__init__.py) does per call: two torch.empty, pointer tables, an 18-argument
// ctypes call, the hop tuples.  Measured end to end that host work is the GPU's idle time between two steps,
// so the same sequence is offered here as ONE compiled function over the SAME C objects: `sampler` is the qv_sampler* the
// ctypes Quiver owns, `table` the address of its ShardTensor's qv_shard_table (0: no gather).  Returns None where the ctypes
// path would raise Unsupported (the caller then takes that path and gets its error), else
// (n_id, [(edge_index, n_src, n_dst[, e_id]) innermost hop first], x or None).
py::object khop_raw(uintptr_t sampler, const torch::Tensor &seeds, const std::vector<int64_t> &sizes, uint64_t rand_seed, int device,
                    int64_t node_count, uintptr_t table, const c10::optional<torch::Tensor> &feature_order, int64_t row_bytes,
                    const std::vector<int64_t> &row_shape, at::ScalarType dtype, int variant, bool with_eid, int64_t max_bytes)
{
    const int n_hops = static_cast<int>(sizes.size());
    if (n_hops < 1 || n_hops > QV_MAX_HOPS) return py::none();
    TORCH_CHECK(seeds.scalar_type() == torch::kInt64, "seeds must be a torch.long tensor");
    TORCH_CHECK(seeds.is_contiguous(), "seeds must be contiguous");
    if (seeds.is_cuda()) {
        TORCH_CHECK(seeds.get_device() == device, "seeds lives on another device than this object");
    } else {  // pinned host memory is device-visible at the same address: hop 0 reads the seeds in place
        TORCH_CHECK(seeds.is_pinned(), "seeds must be a CUDA tensor (or pinned host memory)");
    }
    const int64_t S = seeds.numel();
    int64_t bn[QV_MAX_HOPS + 1], be[QV_MAX_HOPS];
    if (qv_khop_bounds(S, sizes.data(), n_hops, bn, be) != QV_OK) return py::none();
    int64_t total = (std::max<int64_t>(bn[n_hops], 1) + 1) / 2 * 2, offs[QV_MAX_HOPS], eoffs[QV_MAX_HOPS];
    for (int h = 0; h < n_hops; h++) {
        offs[h] = total;
        total += std::max<int64_t>(2 * be[h], 2);
    }
    if (with_eid)
        for (int h = 0; h < n_hops; h++) {
            eoffs[h] = total;
            total += std::max<int64_t>(be[h], 2);
        }
    const auto dev = torch::Device(torch::kCUDA, device);
    int64_t x_rows = 0;
    if (table) {
        TORCH_CHECK(at::cuda::current_device() == device, "sample_khop(gather=...) must run with the sampler's device current");
        x_rows = std::max<int64_t>(std::min(bn[n_hops], node_count + S), 1);
        if (x_rows * row_bytes > max_bytes) return py::none();
    }
    torch::Tensor arena, x;
    int64_t nodes[QV_MAX_HOPS + 1], edges[QV_MAX_HOPS];
    int rc;
    {
        py::gil_scoped_release nogil;
        arena = torch::empty({total}, torch::TensorOptions().dtype(torch::kInt64).device(dev));
        int64_t *base = arena.data_ptr<int64_t>(), *bufs[QV_MAX_HOPS], *eids[QV_MAX_HOPS];
        for (int h = 0; h < n_hops; h++) {
            bufs[h] = base + offs[h];
            eids[h] = with_eid ? base + eoffs[h] : nullptr;
        }
        const int64_t *v = seeds.data_ptr<int64_t>();
        auto *s = reinterpret_cast<qv_sampler *>(sampler);
        if (!table) {
            rc = qv_khop(s, v, S, sizes.data(), n_hops, rand_seed, base, bufs, with_eid ? eids : nullptr, nodes, edges,
                         cur_stream(device));
        } else {
            const int64_t *order = nullptr;
            if (feature_order.has_value() && feature_order->defined()) order = long_ptr(*feature_order, "feature_order", device);
            std::vector<int64_t> shape{x_rows};
            shape.insert(shape.end(), row_shape.begin(), row_shape.end());
            x = torch::empty(shape, torch::TensorOptions().dtype(dtype).device(dev));
            rc = qv_khop_gather(s, v, S, sizes.data(), n_hops, rand_seed, base, bufs, with_eid ? eids : nullptr,
                                reinterpret_cast<const qv_shard_table *>(table), order, row_bytes, x.data_ptr(), x_rows, variant,
                                nodes, edges, cur_stream(device));
        }
    }
    if (rc == QV_ERR_UNSUPPORTED) return py::none();
    ok(rc);
    py::list hops;
    for (int h = 0; h < n_hops; h++) {
        auto ei = arena.narrow(0, offs[h], 2 * edges[h]).view({2, edges[h]});
        if (with_eid)
            hops.append(py::make_tuple(ei, nodes[h + 1], nodes[h], arena.narrow(0, eoffs[h], edges[h])));
        else
            hops.append(py::make_tuple(ei, nodes[h + 1], nodes[h]));
    }
    if (!table) return py::make_tuple(arena.narrow(0, 0, nodes[n_hops]), hops, py::none());
    return py::make_tuple(arena.narrow(0, 0, nodes[n_hops]), hops, x.narrow(0, 0, nodes[n_hops]));
}

// The same for ShardTensor.gather of the ctypes package: `table` = address of its qv_shard_table; returns
// table[feature_order[indices]] on torch's current stream (qv_gather).
torch::Tensor gather_raw(uintptr_t table, const torch::Tensor &indices, const c10::optional<torch::Tensor> &feature_order,
                         int64_t row_bytes, const std::vector<int64_t> &row_shape, at::ScalarType dtype, int variant)
{
    const int device = indices.is_cuda() ? indices.get_device() : -1;
    const int64_t *idx = long_ptr(indices, "indices", -1);
    const int64_t *order = nullptr;
    if (feature_order.has_value() && feature_order->defined()) order = long_ptr(*feature_order, "feature_order", device);
    py::gil_scoped_release nogil;
    c10::cuda::CUDAGuard guard(device);
    std::vector<int64_t> shape{indices.numel()};
    shape.insert(shape.end(), row_shape.begin(), row_shape.end());
    auto out = torch::empty(shape, torch::TensorOptions().dtype(dtype).device(indices.device()));
    ok(qv_gather(reinterpret_cast<const qv_shard_table *>(table), idx, order, indices.numel(), row_bytes, out.data_ptr(), variant,
                 cur_stream(device)));
    return out;
}
}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m)
{
    m.doc() = "torch_quiver: the reference's pybind11 surface on top of libquiver_b200.so (C ABI, sm_100a)";
    m.def("device_quiver_from_csr_array", &device_quiver_from_csr_array, py::arg("indptr"), py::arg("indices"),
          py::arg("edge_ids"), py::arg("device") = 0, py::arg("cuda") = false);
    py::class_<Quiver, std::shared_ptr<Quiver>>(m, "Quiver")
        .def("sample_sub", &Quiver::sample_sub, py::call_guard<py::gil_scoped_release>())
        .def("sample_neighbor", &Quiver::sample_neighbor, py::call_guard<py::gil_scoped_release>())
        .def("cal_neighbor_prob", &Quiver::cal_neighbor_prob, py::call_guard<py::gil_scoped_release>())
        .def("reindex_single", &Quiver::reindex_single, py::call_guard<py::gil_scoped_release>())
        .def("sample_khop", &Quiver::sample_khop, py::call_guard<py::gil_scoped_release>())
        .def_readwrite("rand_seed", &Quiver::rand_seed);
    m.def("init_p2p", [](std::vector<int> devices) {
        int n = 0;
        ok(qv_init_p2p(devices.data(), static_cast<int>(devices.size()), &n));
        return n;
    }, py::call_guard<py::gil_scoped_release>());
    m.def("can_device_access_peer", [](int src, int dst) {
        int r = 0;
        ok(qv_can_device_access_peer(src, dst, &r));
        return r != 0;
    }, py::call_guard<py::gil_scoped_release>());
    py::class_<ShardTensorItem>(m, "ShardTensorItem")
        .def(py::init<>())
        .def("share_ipc", &ShardTensorItem::share_ipc)
        .def("from_ipc", &ShardTensorItem::from_ipc);
    py::class_<ShardTensor>(m, "ShardTensor")
        .def(py::init<int>())
        .def("__getitem__", &ShardTensor::getitem, py::call_guard<py::gil_scoped_release>())
        .def("unregister", &ShardTensor::unregister, py::call_guard<py::gil_scoped_release>())
        .def("shape", &ShardTensor::shape)
        .def("numel", &ShardTensor::numel)
        .def("device", &ShardTensor::device)
        .def("stride", &ShardTensor::stride)
        .def("size", &ShardTensor::size)
        .def("device_count", &ShardTensor::device_count)
        .def("append", &ShardTensor::append, py::call_guard<py::gil_scoped_release>())
        .def("append", &ShardTensor::append_item, py::call_guard<py::gil_scoped_release>())
        .def("share_ipc", &ShardTensor::share_ipc, py::call_guard<py::gil_scoped_release>());
    m.def("abi_version", []() { return qv_abi_version(); });
    m.def("khop_raw", &khop_raw);
    m.def("gather_raw", &gather_raw);
}
