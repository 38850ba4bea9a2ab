// host_parity — parity test of the CUDA path driven from the C++ host (TEST CODE: the only C++ that touches oracle/).
// The same scene is built through host.hpp into libb200rt.so ("rt_") and into the CPU oracle liborc.so ("orc_"),
// rendered with both, and compared with the bars of BASELINE.json north_star: hit IDs >= 99.99 % identical,
// radiance within 1e-3 relative on >= 99.9 % of the pixels, identical ray counts.
//   host_parity <libb200rt.so> <liborc.so> <assets> <config> <width> <height>
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "../../ray_tracing_gallery_b200/host_cpp/host.hpp"

using namespace b200rt_host;

struct Frame {
    std::vector<uint8_t> rgba8;
    std::vector<float> radiance;
    std::vector<uint32_t> ids;
    uint64_t counts[2] = {0, 0};
};

static Frame render(Backend& be, const SceneSetup& s) {
    Frame f;
    size_t px = (size_t)s.width * s.height;
    f.rgba8.resize(px * 4); f.radiance.resize(px * 3); f.ids.resize(px * s.max_segments * 3);
    RtUniforms u = s.uniforms(1);
    RtRenderParams p = s.params();
    RtFrameOutputs o = {f.rgba8.data(), f.radiance.data(), f.ids.data(), f.counts};
    be.check(be.render(be.ctx, &u, &p, &o), "render");
    return f;
}

int main(int argc, char** argv) {
    if (argc < 7) { std::fprintf(stderr, "usage: host_parity <libb200rt.so> <liborc.so> <assets> <config> <w> <h>\n"); return 2; }
    try {
        Backend gpu, orc;
        gpu.open_b200rt(argv[1], 0);
        orc.open(argv[2], "orc_");
        int (*orc_create)(void**) = nullptr;
        orc.bind(orc_create, "orc_create");
        orc_create(&orc.ctx);
        uint32_t w = (uint32_t)std::atoi(argv[5]), h = (uint32_t)std::atoi(argv[6]);
        Host hg(gpu, argv[3]), ho(orc, argv[3]);
        SceneSetup sg = build_scene(hg, argv[4], w, h), so = build_scene(ho, argv[4], w, h);
        Frame a = render(gpu, sg), b = render(orc, so);
        size_t px = (size_t)w * h, same_ids = 0, within = 0, close8 = 0;
        size_t per = (size_t)sg.max_segments * 3;
        for (size_t i = 0; i < px; i++) {
            bool eq = true;
            for (size_t k = 0; k < per; k++) eq = eq && a.ids[i * per + k] == b.ids[i * per + k];
            same_ids += eq;
            bool ok = true;
            for (int c = 0; c < 3; c++) {
                float x = a.radiance[3 * i + c], y = b.radiance[3 * i + c];
                if (std::isfinite(y)) ok = ok && std::fabs(x - y) / std::fmax(std::fabs(y), 1e-3f) <= 1e-3f;
            }
            within += ok;
            int d = 0;
            for (int c = 0; c < 4; c++) d = std::max(d, std::abs((int)a.rgba8[4 * i + c] - (int)b.rgba8[4 * i + c]));
            close8 += d <= 1;
        }
        double id_frac = (double)same_ids / px, rad_frac = (double)within / px, px_frac = (double)close8 / px;
        bool counts_ok = a.counts[0] == b.counts[0] && a.counts[1] == b.counts[1];
        std::printf("{\"config\": \"%s\", \"hit_id_agreement\": %.6f, \"radiance_within_1e-3\": %.6f, \"rgba8_within_1\": %.6f, \"rays\": [%llu, %llu], \"ray_counts_equal\": %s}\n",
                    argv[4], id_frac, rad_frac, px_frac, (unsigned long long)a.counts[0], (unsigned long long)a.counts[1], counts_ok ? "true" : "false");
        gpu.close();
        orc.close();
        return (id_frac >= 0.9999 && rad_frac >= 0.999 && px_frac >= 0.999 && counts_ok) ? 0 : 1;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "host_parity: %s\n", e.what());
        return 3;
    }
}
