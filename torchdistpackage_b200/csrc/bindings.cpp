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

// Grouped GEMM over G = M / grp_rows groups stacked along the rows of C (MoE experts):
//   C[g] = act(op(A)[g] @ op(B)[g] + bias[g]) with the per-group operand windows described by the
//   coordinate offsets (elements) a_m, a_k, b_n, b_k -- see GemmLaunch.  N and K are ONE group's.
void gemm_grouped(const Tensor& a, const Tensor& b, Tensor& c, bool trans_a, bool trans_b, int64_t N,
                  int64_t K, int64_t grp_rows, int64_t a_m, int64_t a_k, int64_t b_n, int64_t b_k,
                  const OptTensor& bias, const OptTensor& aux_in, const OptTensor& aux_out,
                  int64_t act) {
  check_bf16_2d(a, "a");
  check_bf16_2d(b, "b");
  TORCH_CHECK(c.is_cuda() && c.dim() == 2 && c.stride(1) == 1 && c.scalar_type() == at::kBFloat16);
  TORCH_CHECK(c.size(1) == N && grp_rows > 0 && c.size(0) % grp_rows == 0, "gemm_grouped: C shape");
  const int64_t G = c.size(0) / grp_rows;
  // extents of the stacked operands must cover every group's window
  const int64_t a_rows = trans_a ? (a_k ? K * G : K) : (a_m ? grp_rows : c.size(0));
  const int64_t a_cols = trans_a ? (a_m ? grp_rows : c.size(0)) : (a_k ? K * G : K);
  const int64_t b_rows = trans_b ? (b_n ? N * G : N) : (b_k ? K * G : K);
  const int64_t b_cols = trans_b ? (b_k ? K * G : K) : (b_n ? N * G : N);
  TORCH_CHECK(a.size(0) == a_rows && a.size(1) == a_cols, "gemm_grouped: A is [", a.size(0), ", ",
              a.size(1), "], expected [", a_rows, ", ", a_cols, "]");
  TORCH_CHECK(b.size(0) == b_rows && b.size(1) == b_cols, "gemm_grouped: B is [", b.size(0), ", ",
              b.size(1), "], expected [", b_rows, ", ", b_cols, "]");
  c10::cuda::CUDAGuard guard(a.device());
  tdp::GemmLaunch g{};
  g.a = a.data_ptr(); g.b = b.data_ptr();
  g.lda = static_cast<int>(a.stride(0)); g.ldb = static_cast<int>(b.stride(0));
  g.trans_a = trans_a; g.trans_b = trans_b;
  g.M = static_cast<int>(c.size(0)); g.N = static_cast<int>(N); g.K = static_cast<int>(K);
  g.c = c.data_ptr(); g.ldc = static_cast<int>(c.stride(0));
  g.alpha = 1.f;
  g.bias = opt_ptr(bias);
  if (g.bias) TORCH_CHECK(bias->scalar_type() == at::kBFloat16 && bias->numel() == G * N, "bias [G, N]");
  g.aux_in = opt_ptr(aux_in);
  g.aux_out = const_cast<void*>(opt_ptr(aux_out));
  if (g.aux_in) { check_bf16_2d(*aux_in, "aux_in"); g.ld_aux = static_cast<int>(aux_in->stride(0)); }
  if (g.aux_out) { check_bf16_2d(*aux_out, "aux_out"); g.ld_aux = static_cast<int>(aux_out->stride(0)); }
  g.act = static_cast<int>(act);
  g.split_k = 1;
  g.cta_group = 1;
  g.grp_rows = static_cast<int>(grp_rows);
  g.grp_a_m = static_cast<int>(a_m); g.grp_a_k = static_cast<int>(a_k);
  g.grp_b_n = static_cast<int>(b_n); g.grp_b_k = static_cast<int>(b_k);
  g.grp_bias = g.bias ? static_cast<int>(N) : 0;
  const char* err = nullptr;
  count_launch();
  int rc = tdp::launch_gemm_bf16(g, cur_stream(), &err);
  TORCH_CHECK(rc == 0, "tdp gemm_grouped failed (", rc, "): ", err ? err : "");
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
  m.def("gemm_grouped", &gemm_grouped, py::arg("a"), py::arg("b"), py::arg("c"), py::arg("trans_a"),
        py::arg("trans_b"), py::arg("N"), py::arg("K"), py::arg("grp_rows"), py::arg("a_m") = 0,
        py::arg("a_k") = 0, py::arg("b_n") = 0, py::arg("b_k") = 0, py::arg("bias") = py::none(),
        py::arg("aux_in") = py::none(), py::arg("aux_out") = py::none(), py::arg("act") = 0);
  m.def("num_sms", &tdp::gemm_num_sms);
  m.def("launch_count", []() { return g_launches.load(); });
  register_ext(m);
}
