// tutu_adapters.hpp — the reference-side binding: C++ adapters that plug libtutu_b200.so into
// TutuRenderer's own plugin interfaces.  #include it AFTER the reference's Renderer.hpp, inside the
// reference's single translation unit (its headers define non-inline functions and globals, so the
// whole host is one TU; this header sees only the C ABI, never CUDA).
//
//   CudaPathTracing      : IIntegrator          (reference include/IIntegrator.hpp:17-24)
//        integrate(g) = flatten the PPMGenerator (objects in Scene::objList order, the BVHNode tree
//        the reference itself built, textures, camera, bkgcolor, eta), tutu_scene_upload,
//        tutu_render_path(SPP) straight into g->cam.FrameBuffer.rgb (PathTracing.hpp:501,513).
//   CudaBdpt             : IIntegrator          (same interface; reference include/BDPT.hpp:395-674)
//        integrate(g) = flatten + tutu_scene_upload + tutu_render_bdpt(SPP).
//   CudaIntersectStrategy : IIntersectStrategy  (reference include/IIntersectStrategy.h:7-17)
//        UpdateInter = one-ray tutu_trace_closest; the Intersection record is then filled by the
//        hit object's own intersect() (same arithmetic, so the same t).  Parity tool only: a
//        per-ray virtual call is not how the GPU is fed, and isShadowRayBlocked bypasses the
//        strategy anyway (IIntegrator.hpp:135-153), which is why the integrator is replaced too.
//
// Plug-in point: Renderer::Renderer, reference include/Renderer.hpp:38-49.
#pragma once
#include <cstring>
#include <mutex>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "tutu_b200.h"

namespace tutu_adapt {

inline void check(int rc, const TutuCtx* ctx, const char* what) {
  if (rc != TUTU_OK) {
    const char* m = tutu_last_error(ctx);
    throw std::runtime_error(std::string(what) + ": " + (m ? m : "?"));  // the C ABI never throws
  }
}

struct HostScene {
  std::vector<TutuPrim> prims;
  std::vector<TutuMaterial> mats;
  std::vector<TutuBvhNode> nodes;
  std::vector<TutuTexture> tex[4];
  std::vector<std::vector<float>> texdata[4];
  TutuSceneDesc desc;
};

inline TutuMaterial to_abi(const Material& r) {
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

// PPMGenerator -> TutuSceneDesc.  Requires g->scene.initializeBVH() to have run (Renderer.hpp:53).
inline void flatten(PPMGenerator* g, HostScene& h) {
  h = HostScene();
  std::unordered_map<Object*, int32_t> index;
  for (size_t i = 0; i < g->scene.objList.size(); ++i) {
    Object* o = g->scene.objList[i].get();
    index[o] = (int32_t)i;
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
    const TutuMaterial m = to_abi(o->mtlcolor);  // every Object carries its own copy: de-duplicate
    int32_t mi = -1;
    for (size_t k = 0; k < h.mats.size() && mi < 0; ++k)
      if (memcmp(&h.mats[k], &m, sizeof(m)) == 0) mi = (int32_t)k;
    if (mi < 0) {
      mi = (int32_t)h.mats.size();
      h.mats.push_back(m);
    }
    p.material = mi;
    p.tex_active = o->isTextureActivated ? 1 : 0;
    p.tex_diffuse = o->textureIndex;
    p.tex_normal = o->normalMapIndex;
    p.tex_roughness = o->roughnessMapIndex;
    p.tex_metallic = o->metallicMapIndex;
    h.prims.push_back(p);
  }
  // the reference's own tree, pre-order (BVH.hpp:15-23, getNode() :125)
  if (!g->scene.objList.empty()) {
    struct Item {
      BVHNode* n;
      int32_t parent;
      bool is_right;
    };
    std::vector<Item> st;
    st.push_back({g->scene.BVHaccelerator->getNode(), -1, false});
    while (!st.empty()) {
      Item it = st.back();
      st.pop_back();
      const int32_t me = (int32_t)h.nodes.size();
      const bool leaf = !it.n->left && !it.n->right;
      h.nodes.push_back({-1, -1, leaf ? index.at(it.n->obj) : -1});
      if (it.parent >= 0) (it.is_right ? h.nodes[it.parent].right : h.nodes[it.parent].left) = me;
      if (!leaf) {
        st.push_back({it.n->right, me, true});
        st.push_back({it.n->left, me, false});
      }
    }
  }
  std::vector<Texture*>* maps[4] = {&g->diffuseMaps, &g->normalMaps, &g->roughnessMaps, &g->metallicMaps};
  for (int c = 0; c < 4; ++c) {
    h.texdata[c].resize(maps[c]->size());
    for (size_t i = 0; i < maps[c]->size(); ++i) {
      Texture* t = (*maps[c])[i];
      std::vector<float>& d = h.texdata[c][i];
      d.resize(t->rgb.size() * 3);
      for (size_t k = 0; k < t->rgb.size(); ++k)
        d[3 * k] = t->rgb[k].x, d[3 * k + 1] = t->rgb[k].y, d[3 * k + 2] = t->rgb[k].z;
      h.tex[c].push_back({t->width, t->height, d.data()});
    }
  }
  TutuSceneDesc& d = h.desc;
  memset(&d, 0, sizeof(d));
  d.struct_size = sizeof(d);
  d.n_prims = (uint32_t)h.prims.size();
  d.prims = h.prims.data();
  d.n_materials = (uint32_t)h.mats.size();
  d.materials = h.mats.data();
  d.n_bvh_nodes = (uint32_t)h.nodes.size();
  d.bvh_nodes = h.nodes.empty() ? nullptr : h.nodes.data();
  for (int c = 0; c < 4; ++c) {
    d.n_tex[c] = (uint32_t)h.tex[c].size();
    d.tex[c] = h.tex[c].empty() ? nullptr : h.tex[c].data();
  }
  d.camera.eye[0] = g->cam.position.x, d.camera.eye[1] = g->cam.position.y, d.camera.eye[2] = g->cam.position.z;
  d.camera.viewdir[0] = g->viewdir.x, d.camera.viewdir[1] = g->viewdir.y, d.camera.viewdir[2] = g->viewdir.z;
  d.camera.updir[0] = g->updir.x, d.camera.updir[1] = g->updir.y, d.camera.updir[2] = g->updir.z;
  d.camera.hfov_deg = g->cam.hfov;
  d.camera.width = g->width;
  d.camera.height = g->height;
  d.camera.parallel_projection = g->parallel_projection;
  d.bkgcolor[0] = g->bkgcolor.x, d.bkgcolor[1] = g->bkgcolor.y, d.bkgcolor[2] = g->bkgcolor.z;
  d.eta = g->eta;
}

class DeviceScene {  // one TutuCtx + the flattened scene, shared by both adapters
 public:
  explicit DeviceScene(int device = 0) { check(tutu_ctx_create(device, &ctx_), nullptr, "tutu_ctx_create"); }
  ~DeviceScene() { tutu_ctx_destroy(ctx_); }
  DeviceScene(const DeviceScene&) = delete;
  DeviceScene& operator=(const DeviceScene&) = delete;
  void upload(PPMGenerator* g) {
    flatten(g, host_);
    check(tutu_scene_upload(ctx_, &host_.desc), ctx_, "tutu_scene_upload");
  }
  TutuCtx* ctx() const { return ctx_; }
  const HostScene& host() const { return host_; }

 private:
  TutuCtx* ctx_ = nullptr;
  HostScene host_;
};

}  // namespace tutu_adapt

// IIntegrator plug-in: `integrator = new CudaPathTracing(g, interStrategy)` where Renderer.hpp:43
// says `new PathTracing(g, interStrategy)`.
class CudaPathTracing : public IIntegrator {
 public:
  CudaPathTracing(PPMGenerator* g_, IIntersectStrategy* inters, uint64_t seed = 1, int device = 0)
      : dev_(device), seed_(seed) {
    this->g = g_;
    this->interStrategy = inters;  // unused: traversal happens on the device
  }
  void integrate(PPMGenerator* g_) override {
    dev_.upload(g_);  // the light list is implied by the emissive objects (PPMGenerator.hpp:317-324)
    static_assert(sizeof(Vector3f) == 3 * sizeof(float), "FrameBuffer.rgb must be packed floats");
    tutu_adapt::check(tutu_render_path(dev_.ctx(), (uint32_t)SPP, seed_, &g_->cam.FrameBuffer.rgb[0].x), dev_.ctx(),
                      "tutu_render_path");
  }
  TutuCtx* ctx() const { return dev_.ctx(); }

 private:
  tutu_adapt::DeviceScene dev_;
  uint64_t seed_;
};

// IIntegrator plug-in for `integrator bdpt`: `integrator = new CudaBdpt(g, interStrategy)` where
// Renderer.hpp:47 says `new BDPT(g, interStrategy)`.  The reference ADDS into a frame buffer that
// Camera::initialize pre-filled with bkgcolor (Camera.hpp:28, BDPT.hpp:891); tutu_render_bdpt
// returns bkgcolor + contributions, so the buffer is overwritten with the same value.
class CudaBdpt : public IIntegrator {
 public:
  CudaBdpt(PPMGenerator* g_, IIntersectStrategy* inters, uint64_t seed = 1, int device = 0) : dev_(device), seed_(seed) {
    this->g = g_;
    this->interStrategy = inters;  // unused: traversal happens on the device
  }
  void integrate(PPMGenerator* g_) override {
    dev_.upload(g_);
    tutu_adapt::check(tutu_render_bdpt(dev_.ctx(), (uint32_t)SPP, seed_, &g_->cam.FrameBuffer.rgb[0].x), dev_.ctx(),
                      "tutu_render_bdpt");
  }
  TutuCtx* ctx() const { return dev_.ctx(); }

 private:
  tutu_adapt::DeviceScene dev_;
  uint64_t seed_;
};

// IIntersectStrategy plug-in: `interStrategy = new CudaIntersectStrategy()` at Renderer.hpp:38.
class CudaIntersectStrategy : public IIntersectStrategy {
 public:
  explicit CudaIntersectStrategy(int device = 0) : dev_(device) {}
  void UpdateInter(Intersection& inter, Scene& sce, const Vector3f& rayOrig, const Vector3f& rayDir) override {
    std::lock_guard<std::mutex> lock(mu_);  // PathTracing calls this from 20 threads (PathTracing.hpp:394-429)
    if (!uploaded_) throw std::runtime_error("CudaIntersectStrategy: call bind(g) after Scene::initializeBVH()");
    const float ray[TUTU_RAY_FLOATS] = {rayOrig.x, rayOrig.y, rayOrig.z, 0.f, rayDir.x, rayDir.y, rayDir.z, 0.f};
    TutuHit h;
    tutu_adapt::check(tutu_trace_closest(dev_.ctx(), ray, 1, &h), dev_.ctx(), "tutu_trace_closest");
    inter = Intersection();
    if (h.prim >= 0) sce.objList[h.prim]->intersect(rayOrig, rayDir, inter);  // fills pos/Ng/Ns/material
  }
  // The reference never calls getShadowCoeffi; any hit within the segment blocks.
  float getShadowCoeffi(Scene&, Intersection& p, Vector3f& lightpos) override {
    std::lock_guard<std::mutex> lock(mu_);
    Vector3f orig = p.pos + 0.0005f * p.Ng;  // BVHStrategy.hpp:14-19
    Vector3f dir = normalized(lightpos - orig);
    const float dist = (lightpos - orig).norm();
    const float ray[TUTU_RAY_FLOATS] = {orig.x, orig.y, orig.z, 0.f, dir.x, dir.y, dir.z, dist};
    uint8_t blocked = 0;
    tutu_adapt::check(tutu_trace_any(dev_.ctx(), ray, 1, &blocked), dev_.ctx(), "tutu_trace_any");
    return blocked ? 0.f : 1.f;
  }
  void bind(PPMGenerator* g_) {
    dev_.upload(g_);
    uploaded_ = true;
  }
  TutuCtx* ctx() const { return dev_.ctx(); }

 private:
  tutu_adapt::DeviceScene dev_;
  std::mutex mu_;
  bool uploaded_ = false;
};
