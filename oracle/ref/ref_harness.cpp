// ref_harness — drives the UNMODIFIED reference (bobhansky/TutuRenderer, headers under
// $(REF)/include, compiled where they lie) on scenes given as TUTUSCN1 files.
//
// TEST INFRASTRUCTURE ONLY.  Built by oracle/Makefile into oracle/_ref/ (git-ignored).  Used by
// tests/, tools/make_golden.py and bench.py's cpu_baseline / --impl reference legs; never by the
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
//   ref_harness render <scene.tscene> <spp> <out.f32> [mode]
//        PathTracing::integrate (mode "stock": N_THREAD=20 as shipped) or the reference's own
//        sub_render_pt row worker on every host core (mode "rows").
#include <cmath>
#include <math.h>
namespace std {
using ::powf;  // Material.hpp:145 uses std::powf, which libstdc++ 13 does not declare
}

#include "PPMGenerator.hpp"
#include "Sphere.hpp"
#include "Scene.hpp"
#include "Object.hpp"
#include "Renderer.hpp"
#include "OBJ_Loader.h"

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <thread>
#include <unordered_map>

#include "tutu_b200.h"  // scene-file IO only (host_scene.o); no CUDA entry point is linked

namespace {

struct Quiet {  // the reference chats on std::cout
  std::streambuf* old;
  std::ostringstream sink;
  Quiet() : old(std::cout.rdbuf(sink.rdbuf())) {}
  ~Quiet() { std::cout.rdbuf(old); }
};

[[noreturn]] void die(const std::string& m) {
  fprintf(stderr, "ref_harness: %s\n", m.c_str());
  exit(2);
}

std::string write_config(const TutuCamera& c, int integrator_type) {
  char path[] = "/tmp/ref_harness_cfg_XXXXXX";
  int fd = mkstemp(path);
  if (fd < 0) die("mkstemp failed");
  FILE* f = fdopen(fd, "w");
  // placeholders that satisfy the parser; exact floats are assigned afterwards
  fprintf(f, "imsize %d %d\neye 0 0 0\nviewdir 0 0 1\nhfov %d\nupdir 0 1 0\nbkgcolor 0 0 0 1.0\n",
          c.width, c.height, c.hfov_deg);
  if (c.parallel_projection) fprintf(f, "projection parallel\n");
  fprintf(f, "integrator %s\n", integrator_type == 3 ? "bdpt" : "path");
  fclose(f);
  return path;
}

void apply_camera(PPMGenerator& g, const TutuCamera& c, const float bkg[3], float eta) {
  g.width = c.width;
  g.height = c.height;
  g.hfov = c.hfov_deg;
  g.eyePos = Vector3f(c.eye[0], c.eye[1], c.eye[2]);
  g.viewdir = Vector3f(c.viewdir[0], c.viewdir[1], c.viewdir[2]);
  g.updir = Vector3f(c.updir[0], c.updir[1], c.updir[2]);
  g.bkgcolor = Vector3f(bkg[0], bkg[1], bkg[2]);
  g.eta = eta;
  g.parallel_projection = c.parallel_projection;
  g.cam.width = g.width;  // PPMGenerator.hpp:299-305
  g.cam.height = g.height;
  g.cam.hfov = g.hfov;
  g.cam.position = g.eyePos;
  g.cam.fwdDir = g.viewdir;
  g.cam.upDir = g.updir;
  g.cam.initialize(g.bkgcolor);
}

Material to_ref(const TutuMaterial& m) {
  Material r;
  r.diffuse = Vector3f(m.diffuse[0], m.diffuse[1], m.diffuse[2]);
  r.specular = Vector3f(m.specular[0], m.specular[1], m.specular[2]);
  r.emission = Vector3f(m.emission[0], m.emission[1], m.emission[2]);
  r.mType = (MaterialType)m.type;
  r.alpha = m.alpha;
  r.eta = m.eta;
  r.roughness = m.roughness;
  r.metallic = m.metallic;
  return r;
}
TutuMaterial from_ref(const Material& r) {
  TutuMaterial m;
  m.diffuse[0] = r.diffuse.x, m.diffuse[1] = r.diffuse.y, m.diffuse[2] = r.diffuse.z;
  m.specular[0] = r.specular.x, m.specular[1] = r.specular.y, m.specular[2] = r.specular.z;
  m.emission[0] = r.emission.x, m.emission[1] = r.emission.y, m.emission[2] = r.emission.z;
  m.type = (int32_t)r.mType;
  m.alpha = r.alpha;
  m.eta = r.eta;
  m.roughness = r.roughness;
  m.metallic = r.metallic;
  return m;
}

// scene file -> reference objects
void populate(PPMGenerator& g, const TutuSceneDesc& d) {
  for (int c = 0; c < 4; ++c) {
    std::vector<Texture*>* dst = c == 0   ? &g.diffuseMaps
                                 : c == 1 ? &g.normalMaps
                                 : c == 2 ? &g.roughnessMaps
                                          : &g.metallicMaps;
    for (uint32_t i = 0; i < d.n_tex[c]; ++i) {
      Texture* t = new Texture();
      t->width = d.tex[c][i].width;
      t->height = d.tex[c][i].height;
      size_t n = (size_t)t->width * t->height;
      t->rgb.resize(n);
      for (size_t k = 0; k < n; ++k)
        t->rgb[k] = Vector3f(d.tex[c][i].rgb[3 * k], d.tex[c][i].rgb[3 * k + 1], d.tex[c][i].rgb[3 * k + 2]);
      dst->push_back(t);
    }
  }
  for (uint32_t i = 0; i < d.n_prims; ++i) {
    const TutuPrim& p = d.prims[i];
    std::unique_ptr<Object> o;
    if (p.type == TUTU_PRIM_SPHERE) {
      auto s = std::make_unique<Sphere>(p.v[0], p.v[1], p.v[2], p.v[3]);
      s->objectType = OBJTYPE::SPEHRE;
      o = std::move(s);
    } else {
      auto t = std::make_unique<Triangle>();
      t->objectType = OBJTYPE::TRIANGLE;
      t->v0 = Vector3f(p.v[0], p.v[1], p.v[2]);
      t->v1 = Vector3f(p.v[3], p.v[4], p.v[5]);
      t->v2 = Vector3f(p.v[6], p.v[7], p.v[8]);
      t->n0 = Vector3f(p.n[0], p.n[1], p.n[2]);
      t->n1 = Vector3f(p.n[3], p.n[4], p.n[5]);
      t->n2 = Vector3f(p.n[6], p.n[7], p.n[8]);
      t->uv0 = Vector2f(p.uv[0], p.uv[1]);
      t->uv1 = Vector2f(p.uv[2], p.uv[3]);
      t->uv2 = Vector2f(p.uv[4], p.uv[5]);
      o = std::move(t);
    }
    o->mtlcolor = to_ref(d.materials[p.material]);
    o->isTextureActivated = p.tex_active != 0;
    o->textureIndex = p.tex_diffuse;
    o->normalMapIndex = p.tex_normal;
    o->roughnessMapIndex = p.tex_roughness;
    o->metallicMapIndex = p.tex_metallic;
    o->initializeBound();
    g.scene.add(std::move(o));
  }
}

struct Loaded {
  TutuSceneFile* file = nullptr;
  const TutuSceneDesc* desc = nullptr;
  std::unique_ptr<PPMGenerator> g;
};

Loaded load_scene(const char* path, int integrator_type = 0) {
  Loaded L;
  if (tutu_scene_file_load(path, &L.file) != TUTU_OK) die(tutu_last_error(nullptr));
  L.desc = tutu_scene_file_desc(L.file);
  std::string cfg = write_config(L.desc->camera, integrator_type);
  L.g.reset(new PPMGenerator(strdup(cfg.c_str())));  // keeps the pointer (inputName)
  remove(cfg.c_str());
  apply_camera(*L.g, L.desc->camera, L.desc->bkgcolor, L.desc->eta);
  populate(*L.g, *L.desc);
  return L;
}

// reference objects -> scene file (prims in objList order, materials de-duplicated)
struct Exported {
  std::vector<TutuPrim> prims;
  std::vector<TutuMaterial> mats;
  std::vector<TutuBvhNode> nodes;
};

void export_objects(PPMGenerator& g, Exported& e) {
  for (auto& up : g.scene.objList) {
    Object* o = up.get();
    TutuPrim p;
    memset(&p, 0, sizeof(p));
    if (o->objectType == OBJTYPE::SPEHRE) {
      Sphere* s = static_cast<Sphere*>(o);
      p.type = TUTU_PRIM_SPHERE;
      p.v[0] = s->centerPos.x, p.v[1] = s->centerPos.y, p.v[2] = s->centerPos.z, p.v[3] = s->radius;
    } else {
      Triangle* t = static_cast<Triangle*>(o);
      p.type = TUTU_PRIM_TRIANGLE;
      const Vector3f* vs[3] = {&t->v0, &t->v1, &t->v2};
      const Vector3f* ns[3] = {&t->n0, &t->n1, &t->n2};
      const Vector2f* ts[3] = {&t->uv0, &t->uv1, &t->uv2};
      for (int k = 0; k < 3; ++k) {
        p.v[3 * k] = vs[k]->x, p.v[3 * k + 1] = vs[k]->y, p.v[3 * k + 2] = vs[k]->z;
        p.n[3 * k] = ns[k]->x, p.n[3 * k + 1] = ns[k]->y, p.n[3 * k + 2] = ns[k]->z;
        p.uv[2 * k] = ts[k]->x, p.uv[2 * k + 1] = ts[k]->y;
      }
    }
    TutuMaterial m = from_ref(o->mtlcolor);
    int32_t mi = -1;
    for (size_t k = 0; k < e.mats.size(); ++k)
      if (memcmp(&e.mats[k], &m, sizeof(m)) == 0) mi = (int32_t)k;
    if (mi < 0) {
      mi = (int32_t)e.mats.size();
      e.mats.push_back(m);
    }
    p.material = mi;
    p.tex_active = o->isTextureActivated ? 1 : 0;
    p.tex_diffuse = o->textureIndex;
    p.tex_normal = o->normalMapIndex;
    p.tex_roughness = o->roughnessMapIndex;
    p.tex_metallic = o->metallicMapIndex;
    e.prims.push_back(p);
  }
}

void export_tree(PPMGenerator& g, Exported& e) {
  std::unordered_map<Object*, int32_t> index;
  for (size_t i = 0; i < g.scene.objList.size(); ++i) index[g.scene.objList[i].get()] = (int32_t)i;
  // pre-order walk of the reference's pointer tree (BVH.hpp:15-23)
  struct Item {
    BVHNode* n;
    int32_t parent;
    bool is_right;
  };
  std::vector<Item> st;
  st.push_back({g.scene.BVHaccelerator->getNode(), -1, false});
  while (!st.empty()) {
    Item it = st.back();
    st.pop_back();
    int32_t me = (int32_t)e.nodes.size();
    bool leaf = !it.n->left && !it.n->right;
    e.nodes.push_back({-1, -1, leaf ? index.at(it.n->obj) : -1});
    if (it.parent >= 0) (it.is_right ? e.nodes[it.parent].right : e.nodes[it.parent].left) = me;
    if (!leaf) {
      st.push_back({it.n->right, me, true});
      st.push_back({it.n->left, me, false});
    }
  }
}

void save(const Exported& e, const TutuSceneDesc& like, const char* path) {
  TutuSceneDesc d = like;
  d.struct_size = sizeof(d);
  d.prims = e.prims.data();
  d.n_prims = (uint32_t)e.prims.size();
  d.materials = e.mats.data();
  d.n_materials = (uint32_t)e.mats.size();
  d.bvh_nodes = e.nodes.data();
  d.n_bvh_nodes = (uint32_t)e.nodes.size();
  if (tutu_scene_file_save(&d, path) != TUTU_OK) die(tutu_last_error(nullptr));
}

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
int render(const char* scene, int spp, const char* out_path, const char* mode) {
  Loaded L = load_scene(scene);
  PPMGenerator* g = L.g.get();
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
  if (!strcmp(mode, "stock")) {
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

}  // namespace

int main(int argc, char** argv) {
  static_assert(sizeof(Vector3f) == 12, "Vector3f must be 3 packed floats");
  if (argc < 2) die("usage: see the header of oracle/ref/ref_harness.cpp");
  std::string cmd = argv[1];
  if (cmd == "dump-cornell" && argc == 6)
    return dump_driver_scene("cornell", argv[2], atoi(argv[3]), atoi(argv[4]), argv[5]);
  if (cmd == "export-bvh" && argc == 4) return export_bvh(argv[2], argv[3]);
  if (cmd == "trace" && (argc == 6 || argc == 7))
    return trace(argv[2], argv[3], argv[4], argv[5], argc == 7 ? atoi(argv[6]) : 0);
  if (cmd == "render" && (argc == 5 || argc == 6))
    return render(argv[2], atoi(argv[3]), argv[4], argc == 6 ? argv[5] : "stock");
  die("bad command line");
}
