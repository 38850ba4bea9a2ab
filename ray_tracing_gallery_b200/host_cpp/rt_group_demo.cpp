// rt_group_demo — one frame on several GPUs driven from C++ through the C ABI alone (no torch, no MPI, no CUDA runtime in
// this program): the multi-GPU seams of the reference's host — device creation (src/main.rs:157-204) and the per-frame
// scene update (src/scene.rs:167-204) — as rt_group_create / rt_group_update_instances / rt_group_render_*.
//   rt_group_demo [--ranks N] [--config c2|c3|default] [--width W] [--height H] [--frames F] [--lib path] [--assets dir]
//                 [--sharded-build] [--instances N]
// The parent starts one child process per GPU (rank r on CUDA device r).  Rank 0 makes the 128-byte group id and passes
// it on through a file; from there on everything between the ranks is the library's business (NCCL broadcast of the
// instance records, peer-memory frame on rank 0, shared page-locked host frame).  Every frame the animated instance record
// (DefaultScene: lain's transform, instance 2) is broadcast from rank 0 and refitted on every rank, the frame is rendered
// by all ranks through BOTH paths, and rank 0 checks the two assembled frames bit for bit against its own single-GPU
// rendering of the same frame.  Exit code 0 and a JSON line with "ok": true when every frame matched.
#include <sys/wait.h>
#include <unistd.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <thread>
#include <vector>

#include "host.hpp"

using namespace b200rt_host;

struct GroupApi {
    int (*unique_id)(void*) = nullptr;
    int (*create)(void*, int, int, const void*, uint32_t, uint32_t, void**) = nullptr;
    void (*destroy)(void*) = nullptr;
    const char* (*last_error)(const void*) = nullptr;
    uint32_t (*partition)(const void*, RtRenderParams*) = nullptr;
    int (*update_instances)(void*, int, uint32_t, uint32_t, const RtInstance*, uint32_t) = nullptr;
    int (*build_tlas)(void*, int, const RtInstance*, uint32_t, uint32_t) = nullptr;
    int (*render_device)(void*, uint64_t, const RtUniforms*, const RtRenderParams*) = nullptr;
    int (*render_host)(void*, uint64_t, const RtUniforms*, const RtRenderParams*) = nullptr;
    int (*acquire_host)(void*, uint64_t, uint32_t, const uint8_t**, uint64_t*) = nullptr;
    int (*readback)(void*, uint64_t, void*, size_t) = nullptr;
    int (*release)(void*, uint64_t) = nullptr;
    int (*barrier)(void*) = nullptr;
    void bind_all(Backend& be) {
        be.bind(unique_id, "rt_group_unique_id"); be.bind(create, "rt_group_create"); be.bind(destroy, "rt_group_destroy");
        be.bind(last_error, "rt_group_last_error"); be.bind(partition, "rt_group_partition"); be.bind(update_instances, "rt_group_update_instances");
        be.bind(build_tlas, "rt_group_build_tlas");
        be.bind(render_device, "rt_group_render_device"); be.bind(render_host, "rt_group_render_host"); be.bind(acquire_host, "rt_group_acquire_host");
        be.bind(readback, "rt_group_readback"); be.bind(release, "rt_group_release"); be.bind(barrier, "rt_group_barrier");
    }
};

static std::string dir_of(const std::string& path) {
    size_t p = path.find_last_of('/');
    return p == std::string::npos ? "." : path.substr(0, p);
}

static int run_rank(int rank, int ranks, const std::string& lib, const std::string& assets, const std::string& config, uint32_t width, uint32_t height,
                    uint32_t frames, const std::string& id_file, bool sharded_build, uint32_t num_instances) {
    Backend be;
    GroupApi ga;
    void* group = nullptr;
    auto gcheck = [&](int rc, const char* what) {
        if (rc == 0) return;
        const char* msg = ga.last_error ? ga.last_error(group) : nullptr;
        throw std::runtime_error(std::string(what) + " failed (" + std::to_string(rc) + "): " + (msg ? msg : ""));
    };
    try {
        be.open_b200rt(lib, rank);
        ga.bind_all(be);
        Host host(be, assets);
        SceneSetup s = build_scene(host, config, width, height, num_instances);
        // ---- the 128-byte id: rank 0 makes it, the others read it from the file
        unsigned char id[RT_GROUP_ID_BYTES];
        if (rank == 0) {
            gcheck(ga.unique_id(id), "rt_group_unique_id");
            std::string tmp = id_file + ".tmp";
            FILE* f = std::fopen(tmp.c_str(), "wb");
            if (!f || std::fwrite(id, 1, sizeof(id), f) != sizeof(id)) throw std::runtime_error("cannot write " + tmp);
            std::fclose(f);
            std::rename(tmp.c_str(), id_file.c_str());
        } else {
            FILE* f = nullptr;
            for (int tries = 0; tries < 6000 && !(f = std::fopen(id_file.c_str(), "rb")); tries++) std::this_thread::sleep_for(std::chrono::milliseconds(10));
            if (!f || std::fread(id, 1, sizeof(id), f) != sizeof(id)) throw std::runtime_error("cannot read the group id from " + id_file);
            std::fclose(f);
        }
        gcheck(ga.create(be.ctx, ranks, rank, id, s.width, s.height, &group), "rt_group_create");
        RtRenderParams p = s.params();
        const size_t bytes = (size_t)s.width * s.height * 4;
        std::vector<uint8_t> single(bytes), from_device(bytes);
        const bool animated = s.instances.size() > 2;
        uint64_t seq = 0, rays_total = 0;
        uint32_t mismatched_frames = 0;
        if (sharded_build) {
            // build_tlas for the group: rank 0 hands over the records, every rank builds 1/N of the tree, the treelets
            // are exchanged over NVLink.  The frame must not change: compare with the frame of the replicated tree built above.
            std::vector<uint8_t> before(bytes), after(bytes);
            RtUniforms u = s.uniforms(1);
            RtFrameOutputs o = {before.data(), nullptr, nullptr, nullptr, nullptr};
            be.check(be.render(be.ctx, &u, &p, &o), "render");
            gcheck(ga.build_tlas(group, 0, rank == 0 ? s.instances.data() : nullptr, (uint32_t)s.instances.size(), 1u /* force sharded */), "rt_group_build_tlas");
            o.rgba8 = after.data();
            be.check(be.render(be.ctx, &u, &p, &o), "render");
            if (std::memcmp(before.data(), after.data(), bytes) != 0) {
                std::fprintf(stderr, "rank %d: the frame changed with the sharded TLAS\n", rank);
                mismatched_frames++;
            }
        }
        auto t0 = std::chrono::steady_clock::now();
        for (uint32_t k = 0; k < frames; k++) {
            // DefaultScene::update + write_resources for the group: ONE 64-byte record from rank 0 to every rank, then refit
            if (animated) {
                RtInstance rec = s.instances[2];
                if (rank == 0) {  // only the root's bytes matter: the other ranks receive them through the broadcast
                    float a = 0.05f * (float)(k + 1), c = std::cos(a), sn = std::sin(a);
                    RtInstance base = s.instances[2];
                    for (int r = 0; r < 3; r++) {  // transform * rotation_y(a): columns 0 and 2 mix
                        float m0 = base.transform[r * 4 + 0], m2 = base.transform[r * 4 + 2];
                        rec.transform[r * 4 + 0] = m0 * c - m2 * sn;
                        rec.transform[r * 4 + 2] = m0 * sn + m2 * c;
                    }
                }
                gcheck(ga.update_instances(group, 0, 2, 1, rank == 0 ? &rec : nullptr, RT_UPDATE_REFIT), "rt_group_update_instances");
            }
            RtUniforms u = s.uniforms(1 + k);
            // path 1: rows stored into rank 0's device frame over NVLink peer memory
            seq++;
            gcheck(ga.render_device(group, seq, &u, &p), "rt_group_render_device");
            if (rank == 0) {
                gcheck(ga.readback(group, seq, from_device.data(), bytes), "rt_group_readback");
                gcheck(ga.release(group, seq), "rt_group_release");
            }
            // path 2: every rank copies its strips into the shared page-locked host frame
            seq++;
            gcheck(ga.render_host(group, seq, &u, &p), "rt_group_render_host");
            if (rank == 0) {
                const uint8_t* host_frame = nullptr;
                uint64_t counts[2] = {0, 0};
                gcheck(ga.acquire_host(group, seq, 20000, &host_frame, counts), "rt_group_acquire_host");
                // the same frame on this GPU alone
                uint64_t counts1[2] = {0, 0};
                RtFrameOutputs o = {single.data(), nullptr, nullptr, counts1, nullptr};
                be.check(be.render(be.ctx, &u, &p, &o), "render");
                bool same = std::memcmp(single.data(), from_device.data(), bytes) == 0 && std::memcmp(single.data(), host_frame, bytes) == 0 &&
                            counts[0] == counts1[0] && counts[1] == counts1[1];
                if (!same) {
                    mismatched_frames++;
                    size_t d1 = 0, d2 = 0;
                    for (size_t i = 0; i < bytes; i++) { d1 += single[i] != from_device[i]; d2 += single[i] != host_frame[i]; }
                    std::fprintf(stderr, "frame %u: %zu bytes differ on the device path, %zu on the host path; rays %llu+%llu vs %llu+%llu\n", k, d1, d2,
                                 (unsigned long long)counts[0], (unsigned long long)counts[1], (unsigned long long)counts1[0], (unsigned long long)counts1[1]);
                }
                rays_total += counts[0] + counts[1];
                gcheck(ga.release(group, seq), "rt_group_release");
            }
        }
        gcheck(ga.barrier(group), "rt_group_barrier");
        double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (rank == 0)
            std::printf("{\"host\": \"c++\", \"ranks\": %d, \"config\": \"%s\", \"resolution\": \"%ux%u\", \"frames\": %u, \"rays\": %llu, \"seconds\": %.3f, "
                        "\"instance_broadcast\": %s, \"sharded_build\": %s, \"instances\": %zu, \"mismatched_frames\": %u, \"ok\": %s}\n",
                        ranks, config.c_str(), s.width, s.height, frames, (unsigned long long)rays_total, sec, animated ? "true" : "false",
                        sharded_build ? "true" : "false", s.instances.size(), mismatched_frames, mismatched_frames == 0 ? "true" : "false");
        ga.destroy(group);
        be.close();
        return mismatched_frames == 0 ? 0 : 3;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "rt_group_demo rank %d: %s\n", rank, e.what());
        return 1;
    }
}

int main(int argc, char** argv) {
    std::string self = dir_of(argv[0]);
    std::string lib = self + "/../csrc/libb200rt.so", assets = self + "/../../assets", config = "default", id_file;
    uint32_t width = 640, height = 360, frames = 4, num_instances = 0;
    int ranks = 2, rank = -1;
    bool sharded_build = false;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto next = [&]() -> std::string {
            if (i + 1 >= argc) { std::fprintf(stderr, "missing value for %s\n", a.c_str()); std::exit(2); }
            return argv[++i];
        };
        if (a == "--ranks") ranks = std::atoi(next().c_str());
        else if (a == "--rank") rank = std::atoi(next().c_str());
        else if (a == "--id-file") id_file = next();
        else if (a == "--config") config = next();
        else if (a == "--width") width = (uint32_t)std::atoi(next().c_str());
        else if (a == "--height") height = (uint32_t)std::atoi(next().c_str());
        else if (a == "--frames") frames = (uint32_t)std::atoi(next().c_str());
        else if (a == "--lib") lib = next();
        else if (a == "--assets") assets = next();
        else if (a == "--sharded-build") sharded_build = true;
        else if (a == "--instances") num_instances = (uint32_t)std::atoi(next().c_str());
        else { std::fprintf(stderr, "unknown argument %s\n", a.c_str()); return 2; }
    }
    if (rank >= 0) return run_rank(rank, ranks, lib, assets, config, width, height, frames, id_file, sharded_build, num_instances);
    // ---- parent: one fresh process per GPU (exec of this program with --rank r), bounded by an alarm
    char tmpl[] = "/tmp/b200rt_group_id_XXXXXX";
    int fd = mkstemp(tmpl);
    if (fd >= 0) { close(fd); unlink(tmpl); }
    id_file = tmpl;
    std::vector<pid_t> kids;
    for (int r = 0; r < ranks; r++) {
        pid_t pid = fork();
        if (pid == 0) {
            std::vector<std::string> args = {argv[0], "--rank", std::to_string(r), "--ranks", std::to_string(ranks), "--id-file", id_file, "--config", config,
                                             "--width", std::to_string(width), "--height", std::to_string(height), "--frames", std::to_string(frames),
                                             "--lib", lib, "--assets", assets, "--instances", std::to_string(num_instances)};
            if (sharded_build) args.push_back("--sharded-build");
            std::vector<char*> cargs;
            for (auto& s : args) cargs.push_back(const_cast<char*>(s.c_str()));
            cargs.push_back(nullptr);
            alarm(240);  // a rank that hangs is killed, never the box
            execv(argv[0], cargs.data());
            _exit(127);
        }
        kids.push_back(pid);
    }
    int failed = 0;
    for (pid_t pid : kids) {
        int st = 0;
        waitpid(pid, &st, 0);
        if (!WIFEXITED(st) || WEXITSTATUS(st) != 0) failed++;
    }
    unlink(id_file.c_str());
    return failed ? 1 : 0;
}
