// ref_harness — drives the UNMODIFIED reference (bobhansky/TutuRenderer, headers under
// $(REF)/include, compiled where they lie) on scenes given as TUTUSCN1 files.
//
// TEST INFRASTRUCTURE ONLY.  Built by oracle/Makefile into oracle/_ref/ (git-ignored).  Used by
// tests/, tests/tools/make_golden.py and bench.py's cpu_baseline / --impl reference legs; never by the
// product library.
//
// The reference is a single translation unit by construction (its headers define non-inline
// functions and globals, global.hpp:18-20), so everything is #included here.
//
//   ref_harness dump-cornell <model_dir> <W> <H> <out.tscene>
//        scene of src/main_cornellBox.cpp:24-71 through objl::Loader + PPMGenerator::loadObj,
//        exported with the reference-built BVH topology.
//   ref_harness dump-veach <model_dir> <W> <H> <out.tscene>       (src/main_veach_bdpt.cpp:24-86)
//   ref_harness export-bvh <in.tscene> <out.tscene>
//        rebuild the objects from the file, run Scene::initializeBVH, write the tree back.
//   ref_harness trace <scene.tscene> <rays.f32> <closest|any> <out.bin> [threads]
//        getIntersection / hasIntersection (BVH.hpp:145-194) per ray.
//   ref_harness render-config <config.txt> <model_dir> <spp> <out.f32> [mode]
//        the config parsed by the reference's PPMGenerator (inline geometry, materials, P3 textures)
//        plus the Cornell ceiling light, rendered by the reference's PathTracing
//   ref_harness render-xform <model_dir> <W> <H> <spp> <out.f32> <sx> <sy> <sz> <axis> <deg> <tx> <ty> <tz>
//   ref_harness dump-xform   <model_dir> <W> <H> <out.tscene>     <sx> <sy> <sz> <axis> <deg> <tx> <ty> <tz>
//        Cornell shell + veach_glass.obj placed by the reference's own scaleObj / rotateObj / transObj
//        (PPMGenerator.hpp:210-270), rendered by the reference's PathTracing / exported with its BVH
//   ref_harness load-texture <file.ppm> <out.f32>
//        PPMGenerator::loadTexture (ASCII P3, PPMGenerator.hpp:1027-1084): prints width/height, dumps the texels
//   ref_harness ppm <W> <H> <in.f32> <out.ppm>
//        PPMGenerator::generate (gamma 0.78 quantisation + ASCII P3, PPMGenerator.hpp:140-160,804-845)
//   ref_harness postprocess <mode> <W> <H> <in.f32> <out.f32>
//        the reference's Postprocessor (Postprocessor.hpp:29-197) on a linear float image with its own
//        constants (STRENGTH 2, GAUSSIANLOOP 1, KERNELSIZE 10, STDDEV 30, EXPOSURE 1.5): mode "extract"
//        (getEmmisiveTexture), "blur" (getGaussianBlurTexture once), "bloom" (the chain of
//        performPostProcess before the tone map: extract, 1 + GAUSSIANLOOP blurs, add), "hdr"
//        (getHDRtexture) or "full" (performPostProcess itself, compiled with HDR_BLOOM)
//   ref_harness render <scene.tscene> <spp> <out.f32> [mode]
//        PathTracing::integrate (mode "stock": N_THREAD=20 as shipped) or the reference's own
//        sub_render_pt row worker on every host core (mode "rows"); BDPT::integrate (mode "bdpt",
//        BDPT.hpp:395-674) or sub_render_bdpt on every host core (mode "bdpt-rows").
#include "ref_common.hpp"

namespace {
using namespace refh;

// ---- scenes hard-coded in the reference drivers --------------------------------------------
struct ObjSpec {
  const char* file;
  Material mtl;
};

int dump_driver_scene(const char* which, const char* model_dir, int W, int H, const char* out) {
  std::vector<ObjSpec> specs;
  TutuCamera cam;
  memset(&cam, 0, sizeof(cam));
  float bkg[3] = {0, 0, 0};
  float eta = 1.0f;
  int integ = 0;
  if (!strcmp(which, "cornell")) {
    // src/main_cornellBox.cpp:24-71 + configs/config_cornellBox.txt:1-7
    Material white, light, green, red;
    white.mType = LAMBERTIAN;
    white.diffuse = {0.725f, 0.71f, 0.68f};
    light.diffuse = {0.725f, 0.71f, 0.68f};
    light.emission = {47.8348007, 38.5663986, 31.0807991};
    green.mType = LAMBERTIAN;
    green.diffuse = {0.14f, 0.45f, 0.091f};
    red.mType = LAMBERTIAN;
    red.diffuse = {0.63f, 0.065f, 0.05f};
    specs = {{"cornellBox/floor.obj", white}, {"cornellBox/light.obj", light},
             {"cornellBox/right.obj", green}, {"cornellBox/left.obj", red},
             {"cornellBox/tallbox.obj", white}, {"cornellBox/shortbox.obj", white}};
    cam.eye[0] = 278, cam.eye[1] = 273, cam.eye[2] = -800;
    cam.viewdir[2] = 1;
    cam.updir[1] = 1;
    cam.hfov_deg = 40;
  } else if (!strcmp(which, "veach")) {
    // src/main_veach_bdpt.cpp:24-86 + configs/config_veach_bdpt.txt:1-7.  The driver asks for
    // "veach_slight.obj" (:49) while the file is veach_sLight.obj: on the author's case-insensitive
    // file system the spot light loads (img/veach_bdpt_spp512_2.png shows it), so it is loaded here.
    Material room, Llight, sLight, table, glass, tallLamp;
    room.mType = LAMBERTIAN;
    room.diffuse = {0.725f, 0.71f, 0.68f};
    Llight.diffuse = {0.725f, 0.71f, 0.68f};
    Llight.emission = {500.0, 500.0, 500.0};
    Llight.emission = Llight.emission * 0.5;
    sLight.diffuse = {0.725f, 0.71f, 0.68f};
    sLight.emission = {6999.999881f, 5450.000167f, 3630.000055f};
    sLight.emission = sLight.emission * 0.5;
    table.mType = LAMBERTIAN;
    table.diffuse = {0.32962962985, 0.257976263762, 0.150291711092};
    glass.mType = PERFECT_REFRACTIVE;
    glass.eta = 1.5f;
    tallLamp.mType = MICROFACET_R;
    tallLamp.roughness = 0.2775146484375f;
    tallLamp.metallic = 0.5f;
    tallLamp.diffuse = {0.32962962985, 0.257976263762, 0.150291711092};
    specs = {{"veach_bdpt/veach_room.obj", room},         {"veach_bdpt/veach_Llight.obj", Llight},
             {"veach_bdpt/veach_sLight.obj", sLight},     {"veach_bdpt/veach_table.obj", table},
             {"veach_bdpt/veach_glass.obj", glass},       {"veach_bdpt/veach_tallLamp.obj", tallLamp},
             {"veach_bdpt/veach_wallLamp.obj", room}};
    cam.eye[0] = -0.5f, cam.eye[1] = 0, cam.eye[2] = 7.6f;
    cam.viewdir[0] = -0.005f, cam.viewdir[2] = -1;
    cam.updir[1] = 1;
    cam.hfov_deg = 40;
    integ = 3;
  } else {
    die(std::string("unknown driver scene ") + which);
  }
  cam.width = W;
  cam.height = H;
  std::string cfg = write_config(cam, integ);
  PPMGenerator g(strdup(cfg.c_str()));
  remove(cfg.c_str());
  apply_camera(g, cam, bkg, eta);
  {
    Quiet q;
    for (auto& s : specs) {
      objl::Loader loader;
      std::string p = std::string(model_dir) + "/" + s.file;
      if (!loader.LoadFile(p)) die("cannot load " + p);
      g.loadObj(loader, s.mtl, -1, -1);
    }
    g.scene.initializeBVH();
  }
  Exported e;
  export_objects(g, e);
  export_tree(g, e);
  TutuSceneDesc like;
  memset(&like, 0, sizeof(like));
  like.camera = cam;
  memcpy(like.bkgcolor, bkg, 12);
  like.eta = eta;
  save(e, like, out);
  printf("{\"prims\": %zu, \"materials\": %zu, \"nodes\": %zu}\n", e.prims.size(), e.mats.size(),
         e.nodes.size());
  return 0;
}

int export_bvh(const char* in, const char* out) {
  Loaded L = load_scene(in);
  {
    Quiet q;
    L.g->scene.initializeBVH();
  }
  Exported e;
  export_objects(*L.g, e);
  export_tree(*L.g, e);
  // keep the file's own material table / indices (export_objects de-duplicates again)
  save(e, *L.desc, out);
  printf("{\"prims\": %zu, \"nodes\": %zu}\n", e.prims.size(), e.nodes.size());
  return 0;
}

// ---- ray batches ----------------------------------------------------------------------------
struct HitOut {
  int32_t prim;
  float t, u, v;
};

int trace(const char* scene, const char* rays_path, const char* kind, const char* out_path,
          int threads) {
  Loaded L = load_scene(scene);
  {
    Quiet q;
    L.g->scene.initializeBVH();
  }
  std::ifstream rf(rays_path, std::ios::binary | std::ios::ate);
  if (!rf) die("cannot open rays file");
  size_t bytes = (size_t)rf.tellg();
  rf.seekg(0);
  size_t n = bytes / (TUTU_RAY_FLOATS * sizeof(float));
  std::vector<float> rays(n * TUTU_RAY_FLOATS);
  rf.read((char*)rays.data(), (std::streamsize)(n * TUTU_RAY_FLOATS * sizeof(float)));
  std::unordered_map<Object*, int32_t> index;
  for (size_t i = 0; i < L.g->scene.objList.size(); ++i)
    index[L.g->scene.objList[i].get()] = (int32_t)i;
  BVHNode* root = L.g->scene.BVHaccelerator->getNode();
  const bool any = !strcmp(kind, "any");
  std::vector<HitOut> hits(any ? 0 : n);
  std::vector<uint8_t> blocked(any ? n : 0);
  if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
  if (threads <= 0) threads = 1;
  auto t0 = std::chrono::steady_clock::now();
  std::vector<std::thread> pool;
  for (int w = 0; w < threads; ++w)
    pool.emplace_back([&, w] {
      size_t lo = n * w / threads, hi = n * (w + 1) / threads;
      for (size_t i = lo; i < hi; ++i) {
        const float* r = &rays[i * TUTU_RAY_FLOATS];
        Vector3f o(r[0], r[1], r[2]), d(r[4], r[5], r[6]);
        if (any) {
          blocked[i] = hasIntersection(root, o, d, r[7]) ? 1 : 0;
          continue;
        }
        Intersection it = getIntersection(root, o, d);
        HitOut h{-1, it.t, 0.f, 0.f};
        if (it.intersected) {
          h.prim = index.at(it.obj);
          if (it.obj->objectType == OBJTYPE::TRIANGLE) {
            // u,v are locals of Triangle::intersect (Triangle.hpp:25-47); same expressions
            Triangle* tr = static_cast<Triangle*>(it.obj);
            Vector3f E1 = tr->v1 - tr->v0;
            Vector3f E2 = tr->v2 - tr->v0;
            Vector3f S = o - tr->v0;
            Vector3f S1 = crossProduct(d, E2);
            Vector3f S2 = crossProduct(S, E1);
            Vector3f rightVec(S2.dot(E2), S1.dot(S), S2.dot(d));
            float left = 1.0f / S1.dot(E1);
            Vector3f res = left * rightVec;
            h.u = res.y;
            h.v = res.z;
          }
        }
        hits[i] = h;
      }
    });
  for (auto& t : pool) t.join();
  double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  std::ofstream of(out_path, std::ios::binary);
  if (any)
    of.write((const char*)blocked.data(), (std::streamsize)blocked.size());
  else
    of.write((const char*)hits.data(), (std::streamsize)(hits.size() * sizeof(HitOut)));
  printf("{\"rays\": %zu, \"seconds\": %.6f, \"threads\": %d, \"kind\": \"%s\"}\n", n, sec, threads,
         kind);
  return 0;
}

// ---- render ---------------------------------------------------------------------------------
// BDPT::integrate's ray-generation constants (BDPT.hpp:396-418)
struct BdptFrame {
  Vector3f ul, delta_h, delta_v, c_off_h, c_off_v, eyePos;
};
BdptFrame bdpt_frame(PPMGenerator* g) {
  Camera& cam = g->cam;
  Vector3f u = normalized(crossProduct(cam.fwdDir, cam.upDir));
  Vector3f v = normalized(crossProduct(u, cam.fwdDir));
  float d = cam.imagePlaneDist;
  float width_half = fabs(tan(degree2Radians(cam.hfov / 2.f)) * d);
  float aspect_ratio = cam.width / (float)cam.height;
  float height_half = width_half / aspect_ratio;
  Vector3f n = normalized(g->viewdir);
  BdptFrame f;
  f.eyePos = cam.position;
  f.ul = f.eyePos + d * n - width_half * u + height_half * v;
  Vector3f ur = f.eyePos + d * n + width_half * u + height_half * v;
  Vector3f ll = f.eyePos + d * n - width_half * u - height_half * v;
  f.delta_h = Vector3f(0, 0, 0);
  f.delta_v = Vector3f(0, 0, 0);
  if (g->width != 1) f.delta_h = (ur - f.ul) / (g->width - 1);
  if (g->height != 1) f.delta_v = (ll - f.ul) / (g->height - 1);
  f.c_off_h = (ur - f.ul) / (float)(g->width * 2);
  f.c_off_v = (ll - f.ul) / (float)(g->height * 2);
  return f;
}

int render_g(PPMGenerator* g, int spp, const char* out_path, const char* mode);

int render(const char* scene, int spp, const char* out_path, const char* mode) {
  const bool bdpt = !strncmp(mode, "bdpt", 4);
  Loaded L = load_scene(scene, bdpt ? 3 : 0);
  return render_g(L.g.get(), spp, out_path, mode);
}

// config file parsed by the reference itself (+ optional Cornell ceiling light), reference PathTracing
int render_config(const char* config, const char* model_dir, int spp, const char* out_path, const char* mode) {
  std::unique_ptr<PPMGenerator> g = load_config(config, model_dir, true);
  return render_g(g.get(), spp, out_path, mode);
}

int render_g(PPMGenerator* g, int spp, const char* out_path, const char* mode) {
  SPP = spp;  // mutable globals, global.hpp:19-20
  SPP_inv = 1.f / SPP;
  std::unique_ptr<Renderer> r;
  {
    Quiet q;
    r.reset(new Renderer(g));  // BVHStrategy + PathTracing + initializeBVH, Renderer.hpp:35-54
    g->initializeLights();     // Renderer.hpp:64
  }
  int threads = 0;
  double sec = 0;
  if (!strcmp(mode, "bdpt-rows")) {
    // the reference's own row worker sub_render_bdpt (BDPT.hpp:679-900) on every host core
    threads = (int)std::thread::hardware_concurrency();
    if (threads <= 0) threads = 1;
    BdptFrame f = bdpt_frame(g);
    Thread_arg_bdpt arg{&f.ul, &f.delta_v, &f.delta_h, &f.c_off_h, &f.c_off_v, &f.eyePos, g,
                        static_cast<BDPT*>(r->integrator)};
    std::atomic<int> next{0};
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> pool;
    for (int w = 0; w < threads; ++w)
      pool.emplace_back([&, w] {
        for (;;) {
          int y = next.fetch_add(1);
          if (y >= g->height) break;
          sub_render_bdpt(&arg, w, y, y + 1);
        }
      });
    for (auto& t : pool) t.join();
    sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  } else if (!strcmp(mode, "stock") || !strcmp(mode, "bdpt")) {
    threads = N_THREAD;
    Quiet q;
    auto t0 = std::chrono::steady_clock::now();
    r->integrator->integrate(g);
    sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  } else {
    // The reference's own row worker sub_render_pt (PathTracing.hpp:485-516) on every host core,
    // rows handed out dynamically.  Ray-generation constants as PathTracing.hpp:357-391.
    threads = (int)std::thread::hardware_concurrency();
    if (threads <= 0) threads = 1;
    Camera& cam = g->cam;
    Vector3f u = normalized(crossProduct(cam.fwdDir, cam.upDir));
    Vector3f v = normalized(crossProduct(u, cam.fwdDir));
    float d = cam.imagePlaneDist;
    if (g->parallel_projection) d = 4.f;
    float width_half = fabs(tan(degree2Radians(cam.hfov / 2.f)) * d);
    float aspect_ratio = cam.width / (float)cam.height;
    float height_half = width_half / aspect_ratio;
    Vector3f n = normalized(g->viewdir);
    Vector3f eyePos = cam.position;
    Vector3f ul = eyePos + d * n - width_half * u + height_half * v;
    Vector3f ur = eyePos + d * n + width_half * u + height_half * v;
    Vector3f ll = eyePos + d * n - width_half * u - height_half * v;
    Vector3f delta_h(0, 0, 0), delta_v(0, 0, 0);
    if (g->width != 1) delta_h = (ur - ul) / (g->width - 1);
    if (g->height != 1) delta_v = (ll - ul) / (g->height - 1);
    Vector3f c_off_h = (ur - ul) / (float)(g->width * 2);
    Vector3f c_off_v = (ll - ul) / (float)(g->height * 2);
    Thread_arg_pt arg{&ul, &delta_v, &delta_h, &c_off_h, &c_off_v, &eyePos, g,
                      static_cast<PathTracing*>(r->integrator)};
    std::atomic<int> next{0};
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> pool;
    for (int w = 0; w < threads; ++w)
      pool.emplace_back([&, w] {
        for (;;) {
          int y = next.fetch_add(1);
          if (y >= g->height) break;
          sub_render_pt(&arg, w, y, y + 1);
        }
      });
    for (auto& t : pool) t.join();
    sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  }
  std::ofstream of(out_path, std::ios::binary);
  of.write((const char*)g->cam.FrameBuffer.rgb.data(),
           (std::streamsize)(g->cam.FrameBuffer.rgb.size() * sizeof(Vector3f)));
  double paths = (double)g->width * g->height * spp;
  printf("{\"width\": %d, \"height\": %d, \"spp\": %d, \"seconds\": %.6f, \"threads\": %d, "
         "\"mode\": \"%s\", \"mpaths_per_s\": %.6f}\n",
         g->width, g->height, spp, sec, threads, mode, paths / sec * 1e-6);
  return 0;
}

// ---- output stage -----------------------------------------------------------------------------
// PPMGenerator::generate (PPMGenerator.hpp:140-160 -> writeHeader :804, writePixel :812) on a given
// linear float image; the reference writes "<config path>.ppm", which is then moved to `out`.
int ppm(int W, int H, const char* in_path, const char* out_path) {
  TutuCamera cam;
  memset(&cam, 0, sizeof(cam));
  cam.viewdir[2] = 1, cam.updir[1] = 1, cam.hfov_deg = 40, cam.width = W, cam.height = H;
  float bkg[3] = {0, 0, 0};
  std::string cfg = write_config(cam, 0);
  // heap object, never destroyed: ~PPMGenerator after generate() crashes in the -O2 build
  PPMGenerator& g = *new PPMGenerator(strdup(cfg.c_str()));
  remove(cfg.c_str());
  apply_camera(g, cam, bkg, 1.0f);
  std::ifstream f(in_path, std::ios::binary);
  if (!f) die("cannot open the float image");
  f.read((char*)g.cam.FrameBuffer.rgb.data(), (std::streamsize)((size_t)W * H * sizeof(Vector3f)));
  {
    Quiet q;
    g.generate();
  }
  std::string produced = cfg + ".ppm";
  if (rename(produced.c_str(), out_path) != 0) {
    std::ifstream src(produced, std::ios::binary);
    std::ofstream dst(out_path, std::ios::binary);
    dst << src.rdbuf();
    remove(produced.c_str());
  }
  printf("{\"width\": %d, \"height\": %d}\n", W, H);
  return 0;
}

// ---- f-2 authoring paths ----------------------------------------------------------------------
int dump_xform(const char* model_dir, int W, int H, const char* out, const Xform& x) {
  std::unique_ptr<PPMGenerator> g = load_xform_scene(model_dir, W, H, x);
  {
    Quiet q;
    g->scene.initializeBVH();
  }
  Exported e;
  export_objects(*g, e);
  export_tree(*g, e);
  TutuSceneDesc like;
  memset(&like, 0, sizeof(like));
  like.camera.eye[0] = 278, like.camera.eye[1] = 273, like.camera.eye[2] = -800;
  like.camera.viewdir[2] = 1, like.camera.updir[1] = 1, like.camera.hfov_deg = 40;
  like.camera.width = W, like.camera.height = H;
  like.eta = 1.0f;
  save(e, like, out);
  printf("{\"prims\": %zu, \"materials\": %zu, \"nodes\": %zu}\n", e.prims.size(), e.mats.size(), e.nodes.size());
  return 0;
}

int load_texture(const char* file, const char* out) {
  TutuCamera cam;
  memset(&cam, 0, sizeof(cam));
  cam.viewdir[2] = 1, cam.updir[1] = 1, cam.hfov_deg = 40, cam.width = 4, cam.height = 4;
  std::string cfg = write_config(cam, 0);
  PPMGenerator& g = *new PPMGenerator(strdup(cfg.c_str()));
  remove(cfg.c_str());
  {
    Quiet q;
    g.loadTexture(file, g.diffuseMaps);
  }
  if (g.diffuseMaps.empty()) die("loadTexture loaded nothing");
  Texture* t = g.diffuseMaps.back();
  std::ofstream of(out, std::ios::binary);
  of.write((const char*)t->rgb.data(), (std::streamsize)(t->rgb.size() * sizeof(Vector3f)));
  printf("{\"width\": %d, \"height\": %d, \"texels\": %zu}\n", t->width, t->height, t->rgb.size());
  return 0;
}

// ---- Postprocessor ----------------------------------------------------------------------------
int postprocess(const char* mode, int W, int H, const char* in_path, const char* out_path) {
  Texture src;
  src.width = W;
  src.height = H;
  src.rgb.resize((size_t)W * H);
  std::ifstream f(in_path, std::ios::binary);
  if (!f) die("cannot open the float image");
  f.read((char*)src.rgb.data(), (std::streamsize)((size_t)W * H * sizeof(Vector3f)));
  Postprocessor p(&src);
  Texture out;
  auto t0 = std::chrono::steady_clock::now();
  if (!strcmp(mode, "extract")) {
    out = p.getEmmisiveTexture(&src);
  } else if (!strcmp(mode, "blur")) {
    out = p.getGaussianBlurTexture(&src, KERNELSIZE, STDDEV);
  } else if (!strcmp(mode, "hdr")) {
    out = p.getHDRtexture(&src);
  } else if (!strcmp(mode, "bloom")) {
    // Postprocessor.hpp:37-50, statement for statement, without the final tone map
    p.renderTextures[1] = p.getEmmisiveTexture(&p.renderTextures[0]);
    int index = 2;
    p.renderTextures[index] = p.getGaussianBlurTexture(&p.renderTextures[1], KERNELSIZE, STDDEV);
    index = 3;
    for (int i = 0; i < GAUSSIANLOOP; i++) {
      p.renderTextures[index] = p.getGaussianBlurTexture(&p.renderTextures[index == 2 ? 3 : 2], KERNELSIZE, STDDEV);
      index = index == 2 ? 3 : 2;
    }
    index = index == 2 ? 3 : 2;
    out = p.add(&p.renderTextures[0], &p.renderTextures[index]);
  } else if (!strcmp(mode, "full")) {
    out = p.performPostProcess();
  } else {
    die(std::string("unknown postprocess mode ") + mode);
  }
  double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  if (out.width != W || out.height != H || out.rgb.size() != (size_t)W * H) die("postprocess: unexpected output size");
  std::ofstream of(out_path, std::ios::binary);
  of.write((const char*)out.rgb.data(), (std::streamsize)(out.rgb.size() * sizeof(Vector3f)));
  printf("{\"width\": %d, \"height\": %d, \"mode\": \"%s\", \"seconds\": %.6f}\n", W, H, mode, sec);
  return 0;
}

}  // namespace

int main(int argc, char** argv) {
  static_assert(sizeof(Vector3f) == 12, "Vector3f must be 3 packed floats");
  if (argc < 2) die("usage: see the header of oracle/ref/ref_harness.cpp");
  std::string cmd = argv[1];
  if (cmd == "dump-cornell" && argc == 6)
    return dump_driver_scene("cornell", argv[2], atoi(argv[3]), atoi(argv[4]), argv[5]);
  if (cmd == "dump-veach" && argc == 6)
    return dump_driver_scene("veach", argv[2], atoi(argv[3]), atoi(argv[4]), argv[5]);
  if (cmd == "export-bvh" && argc == 4) return export_bvh(argv[2], argv[3]);
  if (cmd == "render-config" && (argc == 6 || argc == 7))
    return render_config(argv[2], argv[3], atoi(argv[4]), argv[5], argc == 7 ? argv[6] : "rows");
  if (cmd == "render-xform" && argc == 15) {
    std::unique_ptr<PPMGenerator> g = load_xform_scene(argv[2], atoi(argv[3]), atoi(argv[4]), parse_xform(argv + 7));
    return render_g(g.get(), atoi(argv[5]), argv[6], "rows");
  }
  if (cmd == "dump-xform" && argc == 14) return dump_xform(argv[2], atoi(argv[3]), atoi(argv[4]), argv[5], parse_xform(argv + 6));
  if (cmd == "load-texture" && argc == 4) return load_texture(argv[2], argv[3]);
  if (cmd == "postprocess" && argc == 7) return postprocess(argv[2], atoi(argv[3]), atoi(argv[4]), argv[5], argv[6]);
  if (cmd == "ppm" && argc == 6) return ppm(atoi(argv[2]), atoi(argv[3]), argv[4], argv[5]);
  if (cmd == "trace" && (argc == 6 || argc == 7))
    return trace(argv[2], argv[3], argv[4], argv[5], argc == 7 ? atoi(argv[6]) : 0);
  if (cmd == "render" && (argc == 5 || argc == 6))
    return render(argv[2], atoi(argv[3]), argv[4], argc == 6 ? argv[5] : "stock");
  die("bad command line");
}
