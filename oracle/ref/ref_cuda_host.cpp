// ref_cuda_host — the UNMODIFIED reference host (PPMGenerator, objl::Loader, Scene, BVHAccel,
// Camera) driving the CUDA core through include/tutu_adapters.hpp, i.e. what a maintainer gets by
// swapping the two `new` expressions in Renderer::Renderer.
//
// TEST INFRASTRUCTURE (integration test of the drop-in boundary): built by oracle/Makefile into
// oracle/_ref/, links ../../tuturenderer_b200/libtutu_b200.so, needs a GPU to run.
//
//   ref_cuda_host render-cornell <model_dir> <W> <H> <spp> <seed> <out.f32>
//        src/main_cornellBox.cpp's scene via objl::Loader + loadObj; CudaPathTracing::integrate
//   ref_cuda_host render-config <config.txt> <model_dir> <spp> <seed> <out.f32>
//        a config file parsed by the reference's PPMGenerator (+ the Cornell ceiling light) through
//        CudaPathTracing::integrate
//   ref_cuda_host render-xform <model_dir> <W> <H> <spp> <seed> <out.f32> <sx> <sy> <sz> <axis> <deg> <tx> <ty> <tz>
//        Cornell shell + veach_glass.obj placed by the reference's own scaleObj / rotateObj / transObj
//        (PPMGenerator.hpp:210-270) before loadObj; the adapter flattens what the host authored
//   ref_cuda_host render <scene.tscene> <spp> <seed> <out.f32>
//   ref_cuda_host trace <scene.tscene> <rays.f32> <out.bin>
//        every ray through CudaIntersectStrategy::UpdateInter and through BVHStrategy; writes the
//        CUDA strategy's {prim,t} and exits non-zero if any Intersection differs from the reference's
#include "ref_common.hpp"
#include "tutu_adapters.hpp"

using namespace refh;

static void dump_fb(PPMGenerator& g, const char* out) {
  std::ofstream of(out, std::ios::binary);
  of.write((const char*)g.cam.FrameBuffer.rgb.data(), (std::streamsize)(g.cam.FrameBuffer.rgb.size() * sizeof(Vector3f)));
}

static int run_integrate(PPMGenerator& g, int spp, uint64_t seed, const char* out) {
  SPP = spp;
  SPP_inv = 1.f / SPP;
  {
    Quiet q;
    g.scene.initializeBVH();  // Renderer.hpp:53
    g.initializeLights();     // Renderer.hpp:64
  }
  BVHStrategy unused;
  CudaPathTracing integrator(&g, &unused, seed);
  auto t0 = std::chrono::steady_clock::now();
  integrator.integrate(&g);  // Renderer.hpp:65
  double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  dump_fb(g, out);
  printf("{\"width\": %d, \"height\": %d, \"spp\": %d, \"seconds\": %.6f, \"mpaths_per_s\": %.3f}\n", g.width, g.height,
         spp, sec, (double)g.width * g.height * spp / sec * 1e-6);
  return 0;
}

int main(int argc, char** argv) {
  try {
    std::string cmd = argc > 1 ? argv[1] : "";
    if (cmd == "render-cornell" && argc == 8) {
      TutuCamera cam;
      memset(&cam, 0, sizeof(cam));
      cam.eye[0] = 278, cam.eye[1] = 273, cam.eye[2] = -800;  // configs/config_cornellBox.txt
      cam.viewdir[2] = 1, cam.updir[1] = 1, cam.hfov_deg = 40;
      cam.width = atoi(argv[3]), cam.height = atoi(argv[4]);
      float bkg[3] = {0, 0, 0};
      std::string cfg = write_config(cam, 0);
      PPMGenerator g(strdup(cfg.c_str()));
      remove(cfg.c_str());
      apply_camera(g, cam, bkg, 1.0f);
      Material white, light, green, red;  // src/main_cornellBox.cpp:24-71
      white.mType = LAMBERTIAN, white.diffuse = {0.725f, 0.71f, 0.68f};
      light.diffuse = {0.725f, 0.71f, 0.68f}, light.emission = {47.8348007, 38.5663986, 31.0807991};
      green.mType = LAMBERTIAN, green.diffuse = {0.14f, 0.45f, 0.091f};
      red.mType = LAMBERTIAN, red.diffuse = {0.63f, 0.065f, 0.05f};
      const std::pair<const char*, Material*> specs[] = {{"floor", &white}, {"light", &light},   {"right", &green},
                                                         {"left", &red},    {"tallbox", &white}, {"shortbox", &white}};
      {
        Quiet q;
        for (auto& s : specs) {
          objl::Loader loader;
          std::string p = std::string(argv[2]) + "/cornellBox/" + s.first + ".obj";
          if (!loader.LoadFile(p)) die("cannot load " + p);
          g.loadObj(loader, *s.second, -1, -1);
        }
      }
      return run_integrate(g, atoi(argv[5]), strtoull(argv[6], nullptr, 10), argv[7]);
    }
    if (cmd == "render-config" && argc == 7) {
      // f-2: the reference's own config parser (inline spheres / triangles, texture state machine,
      // PPMGenerator.hpp:328-482, 584-764) feeds the flattener of include/tutu_adapters.hpp
      std::unique_ptr<PPMGenerator> g = load_config(argv[2], argv[3], true);
      return run_integrate(*g, atoi(argv[4]), strtoull(argv[5], nullptr, 10), argv[6]);
    }
    if (cmd == "render-xform" && argc == 16) {
      std::unique_ptr<PPMGenerator> g = load_xform_scene(argv[2], atoi(argv[3]), atoi(argv[4]), parse_xform(argv + 8));
      return run_integrate(*g, atoi(argv[5]), strtoull(argv[6], nullptr, 10), argv[7]);
    }
    if (cmd == "render" && argc == 6) {
      Loaded L = load_scene(argv[2]);
      return run_integrate(*L.g, atoi(argv[3]), strtoull(argv[4], nullptr, 10), argv[5]);
    }
    if (cmd == "trace" && argc == 5) {
      Loaded L = load_scene(argv[2]);
      {
        Quiet q;
        L.g->scene.initializeBVH();
      }
      std::ifstream rf(argv[3], std::ios::binary | std::ios::ate);
      if (!rf) die("cannot open rays file");
      size_t n = (size_t)rf.tellg() / (TUTU_RAY_FLOATS * sizeof(float));
      rf.seekg(0);
      std::vector<float> rays(n * TUTU_RAY_FLOATS);
      rf.read((char*)rays.data(), (std::streamsize)(rays.size() * sizeof(float)));
      CudaIntersectStrategy cuda;
      cuda.bind(L.g.get());
      BVHStrategy cpu;
      std::unordered_map<Object*, int32_t> index;
      for (size_t i = 0; i < L.g->scene.objList.size(); ++i) index[L.g->scene.objList[i].get()] = (int32_t)i;
      struct Out {
        int32_t prim;
        float t;
      };
      std::vector<Out> out(n);
      size_t bad = 0;
      for (size_t i = 0; i < n; ++i) {
        const float* r = &rays[i * TUTU_RAY_FLOATS];
        Vector3f o(r[0], r[1], r[2]), d(r[4], r[5], r[6]);
        Intersection a, b;
        static_cast<IIntersectStrategy&>(cuda).UpdateInter(a, L.g->scene, o, d);
        static_cast<IIntersectStrategy&>(cpu).UpdateInter(b, L.g->scene, o, d);
        out[i] = {a.intersected ? index.at(a.obj) : -1, a.t};
        if (a.intersected != b.intersected || a.obj != b.obj || memcmp(&a.t, &b.t, 4) || memcmp(&a.pos, &b.pos, 12) ||
            memcmp(&a.Ns, &b.Ns, 12))
          ++bad;
      }
      std::ofstream of(argv[4], std::ios::binary);
      of.write((const char*)out.data(), (std::streamsize)(out.size() * sizeof(Out)));
      printf("{\"rays\": %zu, \"mismatches\": %zu}\n", n, bad);
      return bad ? 1 : 0;
    }
    die("bad command line (see the header of oracle/ref/ref_cuda_host.cpp)");
  } catch (const std::exception& e) {
    fprintf(stderr, "ref_cuda_host: %s\n", e.what());
    return 3;
  }
}
