// ref_common.hpp — shared by oracle/ref/ref_harness.cpp and oracle/ref/ref_cuda_host.cpp: pulls the
// UNMODIFIED reference headers into the single translation unit and converts between reference
// objects and TUTUSCN1 scene files.  TEST INFRASTRUCTURE ONLY (see ref_harness.cpp).
#pragma once
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
// Postprocessor::performPostProcess has no return statement unless one of HDR_ONLY / BLOOM_ONLY /
// HDR_BLOOM is defined (Postprocessor.hpp:29-60); HDR_BLOOM selects the whole chain (bloom, then the
// exposure tone map).  Nothing in the reference's drivers defines any of them (the call is commented out).
#define HDR_BLOOM
#include "Postprocessor.hpp"

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <thread>
#include <unordered_map>

#include "tutu_b200.h"  // scene-file IO only (host_scene.o); no CUDA entry point is linked

namespace refh {

struct Quiet {  // the reference chats on std::cout
  std::ostringstream sink;  // declared (hence constructed) before `old`, whose initialiser uses it
  std::streambuf* old;
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

// A config file parsed by the reference's own PPMGenerator (camera, inline `sphere` / `v` / `f` / `vn` /
// `vt` geometry, `mtlcolor` / MICROFACET_* / PERFECT_* materials, `texture` / `bump` /
// `roughnessTexture` / `metallicTexture` maps: PPMGenerator.hpp:328-482, 488-791).  The config
// format has no emission keyword (the drivers set emission in C++, src/main_cornellBox.cpp:31-33), so
// the Cornell ceiling light can be added the way the drivers add it: objl::Loader + loadObj.
std::unique_ptr<PPMGenerator> load_config(const char* config_path, const char* model_dir, bool add_light) {
  std::unique_ptr<PPMGenerator> g;
  {
    Quiet q;
    g.reset(new PPMGenerator(strdup(config_path)));
    if (add_light) {
      Material light;
      light.diffuse = {0.725f, 0.71f, 0.68f};
      light.emission = {47.8348007, 38.5663986, 31.0807991};
      objl::Loader loader;
      std::string p = std::string(model_dir) + "/cornellBox/light.obj";
      if (!loader.LoadFile(p)) die("cannot load " + p);
      g->loadObj(loader, light, -1, -1);
    }
  }
  return g;
}

// f-2 authoring path: the Cornell shell of src/main_cornellBox.cpp plus model/veach_bdpt/veach_glass.obj
// (1214 smooth-shaded triangles) placed with the reference's OWN object transforms, in the order a driver
// would call them: PPMGenerator::scaleObj, rotateObj, transObj (PPMGenerator.hpp:210-270), then loadObj.
struct Xform {
  float sx, sy, sz;
  int axis;
  float degree;
  float tx, ty, tz;
};
std::unique_ptr<PPMGenerator> load_xform_scene(const char* model_dir, int W, int H, const Xform& x) {
  TutuCamera cam;
  memset(&cam, 0, sizeof(cam));
  cam.eye[0] = 278, cam.eye[1] = 273, cam.eye[2] = -800;  // configs/config_cornellBox.txt
  cam.viewdir[2] = 1, cam.updir[1] = 1, cam.hfov_deg = 40;
  cam.width = W, cam.height = H;
  float bkg[3] = {0, 0, 0};
  std::string cfg = write_config(cam, 0);
  std::unique_ptr<PPMGenerator> g(new PPMGenerator(strdup(cfg.c_str())));
  remove(cfg.c_str());
  apply_camera(*g, cam, bkg, 1.0f);
  Material white, light, green, red, glass;
  white.mType = LAMBERTIAN, white.diffuse = {0.725f, 0.71f, 0.68f};
  light.diffuse = {0.725f, 0.71f, 0.68f}, light.emission = {47.8348007, 38.5663986, 31.0807991};
  green.mType = LAMBERTIAN, green.diffuse = {0.14f, 0.45f, 0.091f};
  red.mType = LAMBERTIAN, red.diffuse = {0.63f, 0.065f, 0.05f};
  glass.mType = MICROFACET_T, glass.eta = 1.5f, glass.roughness = 0.2f;
  const std::pair<const char*, Material*> shell[] = {{"floor", &white}, {"light", &light}, {"right", &green},
                                                     {"left", &red}, {"tallbox", &white}};
  Quiet q;
  for (auto& s : shell) {
    objl::Loader loader;
    std::string p = std::string(model_dir) + "/cornellBox/" + s.first + ".obj";
    if (!loader.LoadFile(p)) die("cannot load " + p);
    g->loadObj(loader, *s.second, -1, -1);
  }
  objl::Loader loader;
  std::string p = std::string(model_dir) + "/veach_bdpt/veach_glass.obj";
  if (!loader.LoadFile(p)) die("cannot load " + p);
  g->scaleObj(loader, x.sx, x.sy, x.sz);
  g->rotateObj(loader, x.axis, x.degree);
  g->transObj(loader, x.tx, x.ty, x.tz);
  g->loadObj(loader, glass, -1, -1);
  return g;
}
inline Xform parse_xform(char** a) {
  return Xform{(float)atof(a[0]), (float)atof(a[1]), (float)atof(a[2]), atoi(a[3]), (float)atof(a[4]),
               (float)atof(a[5]), (float)atof(a[6]), (float)atof(a[7])};
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


}  // namespace refh
