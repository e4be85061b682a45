/*
 * ref_present.cpp — the reference's OWN present-pass fragment shader (assets/shaders/image.frag), compiled by g++ and run on the
 * host, one invocation per output texel.  TEST INFRASTRUCTURE, NOT PRODUCT.  See ref_trace.cpp / glsl_compat.h.
 */
#define REFSHADER_NS refshader_present
#include "glsl_compat.h"
// (namespace refshader_present is open)

#include "../_ref/image.frag.inc"

}  // namespace refshader_present

#include <atomic>
#include <mutex>
#include <thread>
#include <vector>

#include "ref_shader.h"

namespace rp = refshader_present;

namespace {
std::mutex g_bind_mutex;
}

extern "C" int ref_present_render(const uint8_t* rgba8_in, uint32_t in_width, uint32_t in_height, const vrt_denoise_params* params, uint32_t out_width,
                                  uint32_t out_height, uint32_t flags, uint8_t* out, int threads) {
    if (!rgba8_in || !params || !out || !in_width || !in_height || !out_width || !out_height || params->samples < 0) return -1;
    std::lock_guard<std::mutex> lock(g_bind_mutex);
    rp::imageSampler.rgba8 = rgba8_in, rp::imageSampler.width = (int)in_width, rp::imageSampler.height = (int)in_height;
    rp::pushConstant.samples = params->samples;  // GraphicsPipeline.zig:27-39
    rp::pushConstant.distributionBias = params->distribution_bias;
    rp::pushConstant.pixelMultiplier = params->pixel_multiplier;
    rp::pushConstant.inversHueTolerance = params->inverse_hue_tolerance;
    const bool bgra = (flags & VRT_DENOISE_BGRA) != 0u;
    int nt = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
    if (nt < 1) nt = 1;
    std::atomic<uint32_t> next_row{0};
    auto worker = [&]() {
        for (;;) {
            const uint32_t y = next_row.fetch_add(1);
            if (y >= out_height) break;
            for (uint32_t x = 0; x < out_width; x++) {
                rp::inUV = rp::vec2(((float)x + 0.5f) / (float)out_width, ((float)y + 0.5f) / (float)out_height);
                rp::shader_main();  // image.frag:74
                uint8_t* p = out + ((size_t)y * out_width + x) * 4;
                const rp::vec4 c = rp::outColor;
                p[bgra ? 2 : 0] = rp::unorm8(c.x), p[1] = rp::unorm8(c.y), p[bgra ? 0 : 2] = rp::unorm8(c.z), p[3] = rp::unorm8(c.w);
            }
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < nt; t++) pool.emplace_back(worker);
    worker();
    for (auto& th : pool) th.join();
    return 0;
}
