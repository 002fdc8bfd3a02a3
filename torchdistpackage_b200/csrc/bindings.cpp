// pybind11 / torch bindings for the sm_100a kernels.  Only this unit includes torch headers.
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>
#include <torch/extension.h>

#include <atomic>
#include <optional>
#include <vector>

#include "common/tdp_api.h"

namespace {

using torch::Tensor;
using OptTensor = std::optional<Tensor>;

inline cudaStream_t cur_stream() { return c10::cuda::getCurrentCUDAStream().stream(); }

// number of kernels of this extension launched so far (bench.py reports the per-step delta)
std::atomic<int64_t> g_launches{0};
inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

inline const void* opt_ptr(const OptTensor& t) {
  return (t.has_value() && t->defined()) ? t->data_ptr() : nullptr;
}

void check_bf16_2d(const Tensor& t, const char* name) {
  TORCH_CHECK(t.is_cuda(), name, " must be a CUDA tensor");
  TORCH_CHECK(t.scalar_type() == at::kBFloat16, name, " must be bf16");
  TORCH_CHECK(t.dim() == 2, name, " must be 2-D");
  TORCH_CHECK(t.stride(1) == 1, name, " must have a contiguous last dim");
}

// C = act(alpha * op(A) @ op(B) + bias) (+ residual), see GemmLaunch
void gemm(const Tensor& a, const Tensor& b, Tensor& c, bool trans_a, bool trans_b,
          const OptTensor& bias, const OptTensor& residual, const OptTensor& aux_in,
          const OptTensor& aux_out, int64_t act, bool accumulate, double alpha, int64_t block_n,
          int64_t max_ctas, int64_t split_k, int64_t cta_group) {
  check_bf16_2d(a, "a");
  check_bf16_2d(b, "b");
  TORCH_CHECK(c.is_cuda() && c.dim() == 2 && c.stride(1) == 1, "c must be a 2-D CUDA tensor");
  TORCH_CHECK(c.scalar_type() == at::kBFloat16 || c.scalar_type() == at::kFloat, "c: bf16|fp32");
  c10::cuda::CUDAGuard guard(a.device());
  tdp::GemmLaunch g{};
  g.a = a.data_ptr();
  g.b = b.data_ptr();
  g.lda = static_cast<int>(a.stride(0));
  g.ldb = static_cast<int>(b.stride(0));
  g.trans_a = trans_a;
  g.trans_b = trans_b;
  g.M = static_cast<int>(trans_a ? a.size(1) : a.size(0));
  g.K = static_cast<int>(trans_a ? a.size(0) : a.size(1));
  const int64_t kb = trans_b ? b.size(1) : b.size(0);
  g.N = static_cast<int>(trans_b ? b.size(0) : b.size(1));
  TORCH_CHECK(kb == g.K, "gemm: inner dimensions differ (", g.K, " vs ", kb, ")");
  TORCH_CHECK(c.size(0) == g.M && c.size(1) == g.N, "gemm: bad output shape");
  g.c = c.data_ptr();
  g.ldc = static_cast<int>(c.stride(0));
  g.c_fp32 = c.scalar_type() == at::kFloat;
  g.accumulate = accumulate;
  g.alpha = static_cast<float>(alpha);
  g.bias = opt_ptr(bias);
  if (g.bias) TORCH_CHECK(bias->scalar_type() == at::kBFloat16 && bias->numel() == g.N, "bias");
  g.residual = opt_ptr(residual);
  if (g.residual) {
    check_bf16_2d(*residual, "residual");
    TORCH_CHECK(residual->size(0) == g.M && residual->size(1) == g.N, "residual shape");
    g.ld_res = static_cast<int>(residual->stride(0));
  }
  g.aux_in = opt_ptr(aux_in);
  g.aux_out = const_cast<void*>(opt_ptr(aux_out));
  if (g.aux_in) {
    check_bf16_2d(*aux_in, "aux_in");
    g.ld_aux = static_cast<int>(aux_in->stride(0));
  }
  if (g.aux_out) {
    check_bf16_2d(*aux_out, "aux_out");
    TORCH_CHECK(!g.aux_in || aux_out->stride(0) == g.ld_aux, "aux ld mismatch");
    g.ld_aux = static_cast<int>(aux_out->stride(0));
  }
  g.act = static_cast<int>(act);
  g.block_n = static_cast<int>(block_n);
  g.max_ctas = static_cast<int>(max_ctas);
  g.split_k = static_cast<int>(split_k);
  g.cta_group = static_cast<int>(cta_group);
  const char* err = nullptr;
  count_launch();
  int rc = tdp::launch_gemm_bf16(g, cur_stream(), &err);
  TORCH_CHECK(rc == 0, "tdp gemm failed (", rc, "): ", err ? err : "");
}

}  // namespace

#include "bindings_ext.inc"

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.doc() = "torchdistpackage_b200 native sm_100a kernels";
  m.def("gemm", &gemm, py::arg("a"), py::arg("b"), py::arg("c"), py::arg("trans_a") = false,
        py::arg("trans_b") = false, py::arg("bias") = py::none(), py::arg("residual") = py::none(),
        py::arg("aux_in") = py::none(), py::arg("aux_out") = py::none(), py::arg("act") = 0,
        py::arg("accumulate") = false, py::arg("alpha") = 1.0, py::arg("block_n") = 0,
        py::arg("max_ctas") = 0, py::arg("split_k") = 1, py::arg("cta_group") = 0);
  m.def("num_sms", &tdp::gemm_num_sms);
  m.def("launch_count", []() { return g_launches.load(); });
  register_ext(m);
}
