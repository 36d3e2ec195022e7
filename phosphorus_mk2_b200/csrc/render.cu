// render.cu — wavefront path tracer on the device (scene upload, ray generation, shading, NEE,
// integration, film).  Filled in after the ray-query path; for now only the context hook.
#include "ctx.hpp"

namespace phos {
struct RenderState {};
void phos_render_release(phos_ctx* ctx) {
  delete ctx->render;
  ctx->render = nullptr;
}
}  // namespace phos
