// jr_ffi.cc -- XLA FFI handlers over the C ABI of libjr_b200.so (include/jr_b200.h): what a JAX host registers as
// custom calls so that `renderer.pipeline.render` (reference renderer/pipeline.py:470-537) and its reverse mode run
// on the B200 kernels.  One forward and one backward target per built-in shader:
//
//     jr_<shader>_forward_ffi   (operands..., zbuffer, canvas) -> (zbuffer', canvas', tri_id)      [depth: no canvas]
//     jr_<shader>_backward_ffi  (operands..., tri_id, d_zbuffer, d_canvas) -> (one gradient per differentiable operand)
//
// Operand ORDER per shader is the table kOperands below -- ffi/jax_binding.py carries the same table and
// tests/test_ffi_sources.py checks that the two agree.  Batching: an operand with one more leading dimension than
// its base rank is batched (batch_stride = elements per image), otherwise shared (batch_stride 0) -- exactly what
// `ffi_call(..., vmap_method="expand_dims")` produces under jax.vmap (size-1 leading axis = shared).
//
// Build (needs jaxlib's headers; not available in the image this repository was developed in, where the file
// compiles to an empty object -- tests/test_ffi_sources.py syntax-checks it against a stub of the FFI API):
//     g++ -std=c++17 -shared -fPIC -I$(python -c "import jaxlib, os; print(os.path.join(os.path.dirname(jaxlib.__file__), 'include'))") \
//         -I include -I /usr/local/cuda/include ffi/jr_ffi.cc -L jaxrenderer_b200/lib -ljr_b200 -o ffi/libjr_ffi.so
#if defined(JR_FFI_STUB)
#include "xla_ffi_stub.h"      // tests only: a minimal stand-in for the XLA FFI API, for -fsyntax-only
#define JR_FFI_ENABLED 1
#elif defined(__has_include)
#if __has_include("xla/ffi/api/ffi.h")
#include "xla/ffi/api/ffi.h"
#define JR_FFI_ENABLED 1
#endif
#endif

#ifdef JR_FFI_ENABLED
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../include/jr_b200.h"

#ifndef JR_FFI_STUB
#include <cuda_runtime_api.h>
#endif

namespace ffi = xla::ffi;

namespace {

// ---- operand tables: name, base (un-batched) rank, float?, differentiable?
struct Operand { const char* name; int rank; bool is_float; bool diff; };

constexpr Operand kCommon[] = {
    {"world_to_clip", 2, true, true}, {"viewport", 2, true, true}, {"position", 2, true, true}, {"faces", 2, false, false}};
constexpr Operand kGouraud[] = {{"normal", 2, true, true}, {"colour", 2, true, true}, {"light_direction", 1, true, true},
                                {"light_colour", 1, true, true}};
constexpr Operand kGouraudTexture[] = {{"normal", 2, true, true}, {"uv", 2, true, false}, {"light_direction", 1, true, true},
                                       {"light_colour", 1, true, true}, {"texture", 3, true, true}};
constexpr Operand kPhong[] = {{"world_to_eye_norm", 2, true, true}, {"normal", 2, true, true}, {"uv", 2, true, false},
                              {"light_direction", 1, true, true}, {"light_colour", 1, true, true}, {"texture", 3, true, true}};
constexpr Operand kPhongDarboux[] = {{"world_to_eye_norm", 2, true, true}, {"normal", 2, true, true}, {"uv", 2, true, true},
                                     {"light_direction", 1, true, true}, {"light_colour", 1, true, true},
                                     {"texture", 3, true, true}, {"normal_map", 3, true, true},
                                     {"id_to_face", 1, false, false}, {"faces_indices", 2, false, false}};
constexpr Operand kPhongReflection[] = {
    {"world_to_eye_norm", 2, true, true}, {"normal", 2, true, true}, {"uv", 2, true, false},
    {"light_colour", 1, true, true}, {"light_dir_eye", 1, true, true}, {"ambient", 1, true, true},
    {"diffuse", 1, true, true}, {"specular", 1, true, true}, {"texture", 3, true, true},
    {"specular_map", 2, true, true}, {"texture_shape", 2, false, false}, {"texture_index", 1, false, false}};
constexpr Operand kShadow[] = {{"shadow_map", 2, true, false}, {"shadow_strength", 1, true, true},
                               {"shadow_world_to_clip", 2, true, false}, {"shadow_viewport", 2, true, false}};

std::vector<Operand> operands(int shader) {
  std::vector<Operand> v(std::begin(kCommon), std::end(kCommon));
  auto add = [&](const Operand* b, const Operand* e) { v.insert(v.end(), b, e); };
  switch (shader) {
    case JR_DEPTH: break;
    case JR_GOURAUD: add(std::begin(kGouraud), std::end(kGouraud)); break;
    case JR_GOURAUD_TEXTURE: add(std::begin(kGouraudTexture), std::end(kGouraudTexture)); break;
    case JR_PHONG: add(std::begin(kPhong), std::end(kPhong)); break;
    case JR_PHONG_DARBOUX: add(std::begin(kPhongDarboux), std::end(kPhongDarboux)); break;
    case JR_PHONG_REFLECTION: add(std::begin(kPhongReflection), std::end(kPhongReflection)); break;
    case JR_PHONG_REFLECTION_SHADOW:
      add(std::begin(kPhongReflection), std::end(kPhongReflection));
      add(std::begin(kShadow), std::end(kShadow));
      break;
  }
  return v;
}

long long batch_stride(const ffi::AnyBuffer& b, int base_rank) {
  auto d = b.dimensions();
  if ((int)d.size() != base_rank + 1 || d[0] == 1) return 0;   // shared (vmap's size-1 axis = in_axes None)
  long long s = 1;
  for (size_t i = 1; i < d.size(); ++i) s *= d[i];
  return s;
}
int64_t dim_from_end(const ffi::AnyBuffer& b, int k) { auto d = b.dimensions(); return d[d.size() - 1 - k]; }

// fill the array fields of JrRenderArgs from the operand list (in table order)
ffi::Error fill(JrRenderArgs& a, int shader, const std::vector<Operand>& ops, ffi::RemainingArgs& args, size_t first = 0) {
  for (size_t i = 0; i < ops.size(); ++i) {
    auto got = args.get<ffi::AnyBuffer>(first + i);
    if (!got.has_value()) return ffi::Error::InvalidArgument(std::string("missing operand ") + ops[i].name);
    const ffi::AnyBuffer& b = *got;
    const std::string n = ops[i].name;
    const long long bs = batch_stride(b, ops[i].rank);
    JrF32 f{static_cast<const float*>(b.untyped_data()), bs};
    JrI32 q{static_cast<const int32_t*>(b.untyped_data()), bs};
    if (n == "world_to_clip") a.world_to_clip = f;
    else if (n == "viewport") a.viewport = f;
    else if (n == "world_to_eye_norm") a.world_to_eye_norm = f;
    else if (n == "position") { a.position = f; a.n_pos = (int32_t)dim_from_end(b, 1); }
    else if (n == "faces") { a.faces = q; a.T = (int32_t)dim_from_end(b, 1); }
    else if (n == "normal") { a.normal = f; a.n_nrm = (int32_t)dim_from_end(b, 1); }
    else if (n == "uv") { a.uv = f; a.n_uv = (int32_t)dim_from_end(b, 1); }
    else if (n == "colour") a.colour = f;
    else if (n == "light_direction") a.light_direction = f;
    else if (n == "light_colour") a.light_colour = f;
    else if (n == "light_dir_eye") a.light_dir_eye = f;
    else if (n == "ambient") a.ambient = f;
    else if (n == "diffuse") a.diffuse = f;
    else if (n == "specular") a.specular = f;
    else if (n == "texture") { a.texture = f; a.tex_w = (int32_t)dim_from_end(b, 2); a.tex_h = (int32_t)dim_from_end(b, 1); }
    else if (n == "specular_map") { a.specular_map = f; a.spec_w = (int32_t)dim_from_end(b, 1); a.spec_h = (int32_t)dim_from_end(b, 0); }
    else if (n == "normal_map") a.normal_map = f;
    else if (n == "texture_shape") { a.texture_shape = q; a.n_objects = (int32_t)dim_from_end(b, 1); }
    else if (n == "texture_index") { a.texture_index = q; a.n_texidx = (int32_t)dim_from_end(b, 0); }
    else if (n == "id_to_face") a.id_to_face = q;
    else if (n == "faces_indices") { a.faces_indices = q; a.n_faces_indices = (int32_t)dim_from_end(b, 1); }
    else if (n == "shadow_map") { a.shadow_map = f; a.shadow_w = (int32_t)dim_from_end(b, 1); a.shadow_h = (int32_t)dim_from_end(b, 0); }
    else if (n == "shadow_strength") a.shadow_strength = f;
    else if (n == "shadow_world_to_clip") a.shadow_world_to_clip = f;
    else if (n == "shadow_viewport") a.shadow_viewport = f;
  }
  a.shader = shader;
  return ffi::Error::Success();
}

ffi::Error status(int st) {
  return st == JR_OK ? ffi::Error::Success() : ffi::Error::Internal(std::string("libjr_b200: ") + jr_strerror(st));
}

// ------------------------------------------------------------------------------------------ forward
// operands..., zbuffer (B,W,H), [canvas (B,W,H,3)], workspace (u8) -> zbuffer' (aliased), [canvas' (aliased)], tri_id
template <int SHADER>
ffi::Error ForwardImpl(cudaStream_t stream, int32_t texture_offset, ffi::RemainingArgs args, ffi::RemainingRets rets) {
  const auto ops = operands(SHADER);
  const bool has_canvas = SHADER != JR_DEPTH;
  const size_t n_in = ops.size() + (has_canvas ? 2 : 1) + 1;
  if (args.size() != n_in) return ffi::Error::InvalidArgument("wrong number of operands");
  if (rets.size() != (has_canvas ? 3u : 2u)) return ffi::Error::InvalidArgument("wrong number of results");
  JrRenderArgs a;
  std::memset(&a, 0, sizeof(a));
  if (auto e = fill(a, SHADER, ops, args); e.failure()) return e;
  a.texture_offset = texture_offset;
  auto z_in = *args.get<ffi::AnyBuffer>(ops.size());
  auto zd = z_in.dimensions();
  if (zd.size() != 3) return ffi::Error::InvalidArgument("zbuffer must be (B, W, H)");
  a.B = (int32_t)zd[0]; a.W = (int32_t)zd[1]; a.H = (int32_t)zd[2];
  auto z_out = *rets.get<ffi::AnyBuffer>(0);
  // zbuffer / canvas are donated (input_output_aliases); if XLA could not alias, copy the incoming values first
  if (z_out->untyped_data() != z_in.untyped_data())
    cudaMemcpyAsync(z_out->untyped_data(), z_in.untyped_data(), z_in.size_bytes(), cudaMemcpyDeviceToDevice, stream);
  a.zbuffer = static_cast<float*>(z_out->untyped_data());
  if (has_canvas) {
    auto c_in = *args.get<ffi::AnyBuffer>(ops.size() + 1);
    auto c_out = *rets.get<ffi::AnyBuffer>(1);
    if (c_out->untyped_data() != c_in.untyped_data())
      cudaMemcpyAsync(c_out->untyped_data(), c_in.untyped_data(), c_in.size_bytes(), cudaMemcpyDeviceToDevice, stream);
    a.canvas = static_cast<float*>(c_out->untyped_data());
  }
  a.tri_id = static_cast<int32_t*>((*rets.get<ffi::AnyBuffer>(has_canvas ? 2 : 1))->untyped_data());
  auto ws = *args.get<ffi::AnyBuffer>(n_in - 1);       // scratch allocated by XLA: jr_workspace_bytes(args) bytes
  a.workspace = ws.untyped_data();
  a.workspace_bytes = ws.size_bytes();
  if (jr_workspace_bytes(&a) > a.workspace_bytes) return status(JR_ERR_WORKSPACE);
  return status(jr_render_forward(&a, stream));
}

// ------------------------------------------------------------------------------------------ backward
// operands..., tri_id, d_zbuffer, [d_canvas], workspace -> one gradient per differentiable operand (table order),
// then d_zbuffer_in, [d_canvas_in] (the cotangents of the incoming buffers)
template <int SHADER>
ffi::Error BackwardImpl(cudaStream_t stream, int32_t texture_offset, ffi::RemainingArgs args, ffi::RemainingRets rets) {
  const auto ops = operands(SHADER);
  const bool has_canvas = SHADER != JR_DEPTH;
  const size_t n_in = ops.size() + 1 + (has_canvas ? 2 : 1) + 1;
  if (args.size() != n_in) return ffi::Error::InvalidArgument("wrong number of operands");
  JrRenderArgs a;
  JrGradArgs g;
  std::memset(&a, 0, sizeof(a));
  std::memset(&g, 0, sizeof(g));
  if (auto e = fill(a, SHADER, ops, args); e.failure()) return e;
  a.texture_offset = texture_offset;
  auto tri = *args.get<ffi::AnyBuffer>(ops.size());
  auto td = tri.dimensions();
  a.B = (int32_t)td[0]; a.W = (int32_t)td[1]; a.H = (int32_t)td[2];
  a.tri_id = static_cast<int32_t*>(tri.untyped_data());
  a.zbuffer = reinterpret_cast<float*>(a.tri_id);   // unused by backward, must be non-NULL
  size_t r = 0;
  for (const Operand& o : ops) {
    if (!o.diff) continue;
    auto out = *rets.get<ffi::AnyBuffer>(r++);
    cudaMemsetAsync(out->untyped_data(), 0, out->size_bytes(), stream);   // accumulated into
    JrF32Out f{static_cast<float*>(out->untyped_data()), 0};
    {
      auto d = out->dimensions();
      if ((int)d.size() == o.rank + 1 && d[0] != 1) { f.batch_stride = 1; for (size_t i = 1; i < d.size(); ++i) f.batch_stride *= d[i]; }
    }
    const std::string n = o.name;
    if (n == "world_to_clip") g.d_world_to_clip = f;
    else if (n == "viewport") g.d_viewport = f;
    else if (n == "world_to_eye_norm") g.d_world_to_eye_norm = f;
    else if (n == "position") g.d_position = f;
    else if (n == "normal") g.d_normal = f;
    else if (n == "colour") g.d_colour = f;
    else if (n == "uv") g.d_uv = f;
    else if (n == "light_direction") g.d_light_direction = f;
    else if (n == "light_colour") g.d_light_colour = f;
    else if (n == "light_dir_eye") g.d_light_dir_eye = f;
    else if (n == "ambient") g.d_ambient = f;
    else if (n == "diffuse") g.d_diffuse = f;
    else if (n == "specular") g.d_specular = f;
    else if (n == "texture") g.d_texture = f;
    else if (n == "specular_map") g.d_specular_map = f;
    else if (n == "normal_map") g.d_normal_map = f;
    else if (n == "shadow_strength") g.d_shadow_strength = f;
  }
  // cotangents of the incoming buffers: copy the output cotangents, the kernels mask them in place (1 - keep)
  auto dz_in = *args.get<ffi::AnyBuffer>(ops.size() + 1);
  auto dz_out = *rets.get<ffi::AnyBuffer>(r++);
  cudaMemcpyAsync(dz_out->untyped_data(), dz_in.untyped_data(), dz_in.size_bytes(), cudaMemcpyDeviceToDevice, stream);
  g.d_zbuffer = static_cast<float*>(dz_out->untyped_data());
  if (has_canvas) {
    auto dc_in = *args.get<ffi::AnyBuffer>(ops.size() + 2);
    auto dc_out = *rets.get<ffi::AnyBuffer>(r++);
    cudaMemcpyAsync(dc_out->untyped_data(), dc_in.untyped_data(), dc_in.size_bytes(), cudaMemcpyDeviceToDevice, stream);
    g.d_canvas = static_cast<float*>(dc_out->untyped_data());
  }
  auto ws = *args.get<ffi::AnyBuffer>(n_in - 1);
  g.workspace = ws.untyped_data();
  g.workspace_bytes = ws.size_bytes();
  if (jr_backward_workspace_bytes(&a, &g) > g.workspace_bytes) return status(JR_ERR_WORKSPACE);
  return status(jr_render_backward(&a, &g, stream));
}

}  // namespace

#define JR_FFI_BIND() \
  ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Attr<int32_t>("texture_offset").RemainingArgs().RemainingRets()
#define JR_FFI_TARGETS(name, id)                                                         \
  XLA_FFI_DEFINE_HANDLER_SYMBOL(jr_##name##_forward_ffi, ForwardImpl<id>, JR_FFI_BIND()); \
  XLA_FFI_DEFINE_HANDLER_SYMBOL(jr_##name##_backward_ffi, BackwardImpl<id>, JR_FFI_BIND());

JR_FFI_TARGETS(depth, JR_DEPTH)
JR_FFI_TARGETS(gouraud, JR_GOURAUD)
JR_FFI_TARGETS(gouraud_texture, JR_GOURAUD_TEXTURE)
JR_FFI_TARGETS(phong, JR_PHONG)
JR_FFI_TARGETS(phong_darboux, JR_PHONG_DARBOUX)
JR_FFI_TARGETS(phong_reflection, JR_PHONG_REFLECTION)
JR_FFI_TARGETS(phong_reflection_shadow, JR_PHONG_REFLECTION_SHADOW)

#endif  // JR_FFI_ENABLED
