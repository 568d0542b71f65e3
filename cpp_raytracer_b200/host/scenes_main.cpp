// scenes_main.cpp -- command-line front end over the host mirror API: builds one of the
// reference's scenes (scenes.hpp) and either dumps it in the flat .scene format (so tests can
// compare it byte for byte with a dump of the reference-built scene) or renders it on the GPU
// through Camera::render -> libb200rt.so and writes a P3 PPM like the reference's main().
//
//   b200rt_scenes <scene> [--w W] [--h H] [--spp S] [--depth D] dump <out.scene>
//   b200rt_scenes <scene> [...] [--gpus N] render <out.ppm>      (N GPUs: samples split inside libb200rt.so)
#include <cstdio>
#include <cstring>
#include <fstream>

#include "scenes.hpp"

int main(int argc, char **argv) {
    if (argc < 3) {
        std::fprintf(stderr, "usage: b200rt_scenes <scene> [--w W --h H --spp S --depth D --gpus N] (dump <file> | render <file.ppm>)\n");
        return 1;
    }
    std::streambuf *cout_buf = std::cout.rdbuf();
    std::ofstream devnull("/dev/null");
    std::cout.rdbuf(devnull.rdbuf());   // keep stdout clean for the caller; progress goes nowhere
    b200rt_scenes::Built s;
    if (!b200rt_scenes::build(argv[1], s)) { std::fprintf(stderr, "unknown scene %s\n", argv[1]); return 1; }
    int i = 2;
    size_t w = 0, h = 0;
    while (i + 1 < argc && std::strncmp(argv[i], "--", 2) == 0) {
        const std::string o = argv[i];
        const unsigned long long v = std::stoull(argv[i + 1]);
        if (o == "--w") w = v;
        else if (o == "--h") h = v;
        else if (o == "--spp") s.camera.set_samples_per_pixel(v);
        else if (o == "--depth") s.camera.set_max_depth(v);
        else if (o == "--gpus") s.camera.set_device_count((int)v);
        else { std::fprintf(stderr, "unknown option %s\n", o.c_str()); return 1; }
        i += 2;
    }
    if (w) s.camera.set_image_width(w);
    if (h) s.camera.set_image_height(h);
    if (i + 1 >= argc) { std::fprintf(stderr, "missing command\n"); return 1; }
    const std::string cmd = argv[i], path = argv[i + 1];
    if (cmd == "dump") {
        b200rt_host::FlatScene flat;
        std::string err;
        if (!b200rt_host::flatten(s.world, flat, err)) { std::fprintf(stderr, "%s\n", err.c_str()); return 2; }
        const B200rtCamera cam = s.camera.to_abi();
        std::ofstream out(path, std::ios::binary);
        const uint64_t hdr[3] = {flat.materials.size(), flat.spheres.size(), flat.quads.size()};
        out.write("B2RTSCN1", 8);
        out.write((const char *)hdr, sizeof hdr);
        out.write((const char *)&cam, sizeof cam);
        out.write((const char *)flat.materials.data(), flat.materials.size() * sizeof(B200rtMaterial));
        out.write((const char *)flat.spheres.data(), flat.spheres.size() * sizeof(B200rtSphere));
        out.write((const char *)flat.quads.data(), flat.quads.size() * sizeof(B200rtQuad));
        std::printf("{\"cmd\":\"dump\",\"materials\":%zu,\"spheres\":%zu,\"quads\":%zu}\n", flat.materials.size(),
                    flat.spheres.size(), flat.quads.size());
    } else if (cmd == "render") {
        s.camera.render(s.world).send_as_ppm(path);
        const B200rtStats &st = s.camera.stats();
        std::printf("{\"cmd\":\"render\",\"paths\":%llu,\"rays\":%llu,\"kernel_ms\":%.3f,\"total_ms\":%.3f,\"build_ms\":%.3f,"
                    "\"devices\":%u,\"replicate_ms\":%.3f,\"exchange_ms\":%.3f,\"peer_exchange\":%u}\n",
                    (unsigned long long)st.paths, (unsigned long long)st.rays, st.kernel_ms, st.total_ms, s.camera.scene_info().build_ms,
                    st.n_devices, st.replicate_ms, st.exchange_ms, st.peer_exchange);
    } else {
        std::fprintf(stderr, "unknown command %s\n", cmd.c_str());
        return 1;
    }
    std::cout.rdbuf(cout_buf);
    return 0;
}
