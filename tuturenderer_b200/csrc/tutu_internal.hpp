// Internal host-side declarations shared by host_scene.cpp (pure C++, no CUDA) and tutu_b200.cu.
// Not part of the ABI.
#pragma once
#include <cstdint>
#include <string>
#include <memory>
#include <type_traits>
#include <utility>
#include <vector>

#include "../../include/tutu_b200.h"

namespace tutu {

// Leaf reference encoding inside a device inner node: ref >= 0 -> inner node index;
// ref < 0 -> ~ref = slot | (sphere ? SPHERE_BIT : 0), slot = leaf number in DFS (pre-order) order.
constexpr uint32_t SPHERE_BIT = 1u << 30;
constexpr uint32_t SLOT_MASK = SPHERE_BIT - 1u;

struct Box {
  float lo[3];
  float hi[3];
};

// 64-byte inner node: both child boxes + child refs (one fetch decides both children).
struct alignas(16) InnerNode {
  float box[12];  // left child: {x.lo, x.hi, y.lo, y.hi, z.lo, z.hi}, then the right child likewise
  int32_t left;
  int32_t right;
  int32_t pad0;
  int32_t pad1;
};
static_assert(sizeof(InnerNode) == 64, "InnerNode must be 64 bytes");

// 48-byte intersection record per leaf slot.
// triangle: {v0.xyz, E1.x} {E1.y, E1.z, E2.x, E2.y} {E2.z, n.xyz}; E1 = v1-v0, E2 = v2-v0,
//           n = normalized(E1 x E2) evaluated with the reference's float expressions.
// sphere:   {c.xyz, r} {0...}
struct alignas(16) LeafGeom {
  float f[12];
};
static_assert(sizeof(LeafGeom) == 48, "LeafGeom must be 48 bytes");

// Compressed 8-wide node of the traversal tree for regular rays (host_wide.cpp builds it, wide.cuh walks it).
// 96 bytes = three 32-byte sectors = three LDG.E.256:
//   [ 0,32)  base2.xyz, scale.xyz, child_base, leaf_base
//   [32,64)  imask, lmask, 6 pad bytes, qlo[x][8], qlo[y][8], qlo[z][8]
//   [64,96)  qhi[x][8], qhi[y][8], qhi[z][8], 8 pad bytes
// Plane decode (bit-identical on host and device): dec(q) = fma(as_float(0x4B000000 | q), scale, base2).
// Slot s holds an inner child (imask bit s; node child_base + rank of s in imask), a leaf (lmask bit s; leaf record
// leaf_base + rank of s in lmask) or nothing (inverted box qlo = 255 > qhi = 0).
struct alignas(32) WideNode {
  float base2[3];
  float scale[3];
  uint32_t child_base;
  uint32_t leaf_base;
  uint8_t imask;
  uint8_t lmask;
  uint8_t pad0[6];
  uint8_t qlo[3][8];
  uint8_t qhi[3][8];
  uint8_t pad1[8];
};
static_assert(sizeof(WideNode) == 96, "WideNode must be 96 bytes");

// Leaf record of the wide tree, in the order the wide nodes reference them: the LeafGeom of the primitive plus
// its leaf code (DFS slot | sphere bit: the tie rule and the result are in terms of the reference's DFS slots).
struct alignas(32) WideLeaf {
  float f[12];
  uint32_t code;
  uint32_t pad[3];
};
static_assert(sizeof(WideLeaf) == 64, "WideLeaf must be 64 bytes");
// exact leaf box (the reference tests it before the primitive, BVH.hpp:148,173), same order as WideLeaf
struct alignas(32) WideLeafBox {
  float lo[3];
  float hi[3];
  float pad[2];
};
static_assert(sizeof(WideLeafBox) == 32, "WideLeafBox must be 32 bytes");

// 64-byte shading record per leaf slot:
// {n0.xyz, uv0.x} {n1.xyz, uv0.y} {n2.xyz, uv1.x} {uv1.y, uv2.x, uv2.y, bits(flags)}
// flags: material index | TEX_ACTIVE_BIT | SPHERE flag bit 30
struct alignas(16) LeafShade {
  float f[15];
  uint32_t flags;
};
static_assert(sizeof(LeafShade) == 64, "LeafShade must be 64 bytes");
constexpr uint32_t TEX_ACTIVE_BIT = 1u << 31;
constexpr uint32_t SHADE_SPHERE_BIT = 1u << 30;
constexpr uint32_t MAT_MASK = (1u << 30) - 1u;

struct alignas(16) LeafTex {
  int32_t diffuse, normal, roughness, metallic;
};

// 64-byte material: {diffuse.xyz, type} {specular.xyz, alpha} {emission.xyz, eta}
//                   {roughness, metallic, has_emission, 0}
struct alignas(16) DevMaterial {
  float diffuse[3];
  int32_t type;
  float specular[3];
  float alpha;
  float emission[3];
  float eta;
  float roughness;
  float metallic;
  int32_t has_emission;
  int32_t pad;
};
static_assert(sizeof(DevMaterial) == 64, "DevMaterial must be 64 bytes");

// Light table entry (PPMGenerator.hpp:317-324 order), 128 bytes.
struct alignas(16) DevLight {
  float v0[3];
  float area;  // Object::getArea()
  float v1[3];
  int32_t type;  // TUTU_PRIM_*
  float v2[3];
  int32_t slot;
  float n0[3];
  float radius;
  float n1[3];
  int32_t material;
  float n2[3];
  int32_t pad0;
  float emission[3];
  int32_t pad1;
  float pad2[4];
};
static_assert(sizeof(DevLight) == 128, "DevLight must be 128 bytes");

struct TexHeader {
  int32_t width, height;
  uint32_t offset;  // in float4 texels into the texel pool
  uint32_t n_texels;
};

// Ray-generation constants, PathTracing.hpp:357-391 + :503.
struct RayGen {
  float eye[3];
  float ul[3];
  float delta_h[3];
  float delta_v[3];
  float c_off_h[3];
  float c_off_v[3];
  int32_t width, height;
};

// BDPT camera constants: Camera.hpp:81-97 after initialize() + the pixel grid of BDPT.hpp:396-418
struct BdptCamConsts {
  float eye[3], fwd[3], ul[3], dh[3], dv[3], coh[3], cov[3];
  float w2r[16];  // world2Raster, row major
  float imagePlaneDist, filmPlaneAreaInv, lensAreaInv;
  int32_t width, height;
};

// std::vector without value-initialisation of trivially constructible elements: resize(n) of the 10^6-entry leaf
// tables would otherwise zero-fill ~200 MB single-threaded before the (parallel) loops that fill them.
template <class T>
struct DefaultInitAllocator : std::allocator<T> {
  template <class U>
  struct rebind {
    using other = DefaultInitAllocator<U>;
  };
  using std::allocator<T>::allocator;
  template <class U>
  void construct(U* p) noexcept(std::is_nothrow_default_constructible<U>::value) {
    ::new (static_cast<void*>(p)) U;
  }
  template <class U, class... A>
  void construct(U* p, A&&... a) {
    ::new (static_cast<void*>(p)) U(std::forward<A>(a)...);
  }
};
template <class T>
using PodVec = std::vector<T, DefaultInitAllocator<T>>;

struct FlatScene {
  uint32_t n_prims = 0;
  uint32_t n_ref_nodes = 0;
  uint32_t depth = 0;
  int32_t root_ref = 0;  // inner index 0, or ~slot when the tree is a single leaf
  bool empty = true;
  Box root_box{};
  float max_edge = 0.f;  // longest triangle edge / sphere diameter (pruning slack)
  PodVec<InnerNode> inner;       // the reference's topology (irregular rays, literal walk)
  PodVec<InnerNode> inner_fast;  // SAH topology over the same leaves (regular rays), host_scene.cpp
  int32_t root_ref_fast = 0;
  uint32_t depth_fast = 0;
  std::vector<WideNode> wide;      // compressed 8-wide collapse of inner_fast (empty: the device walks inner_fast)
  std::vector<WideLeaf> wleaf;
  std::vector<WideLeafBox> wbox;
  uint32_t wide_depth = 0;         // levels of wide nodes, root = 1
  PodVec<LeafGeom> geom;
  PodVec<LeafShade> shade;
  std::vector<LeafTex> leaftex;
  PodVec<int32_t> slot_to_prim;
  PodVec<Box> leaf_box;  // per leaf slot (small-scene fast path, traversal-tree builders)
  PodVec<uint32_t> leaf_code;  // per leaf slot: slot | SPHERE_BIT
  std::vector<DevMaterial> materials;
  std::vector<DevLight> lights;
  std::vector<TexHeader> tex_headers[4];
  std::vector<float> texels;  // float4 per texel (rgb + pad), all channels pooled
  RayGen raygen{};
  BdptCamConsts bdpt_cam{};
  TutuCamera camera{};
  float bkgcolor[3] = {0, 0, 0};
  float eta = 1.f;
};

// Gaussian taps of Postprocessor::getGaussianBlurTexture (Postprocessor.hpp:77-79, 86-97), evaluated on the
// host with the reference's own float expression.  g must hold kernel_size floats.
void post_gaussian_weights(int kernel_size, float stddev, float* g, float* sum, int* start);

// Error plumbing (thread-local message, mirrored into the ctx by the callers in tutu_b200.cu).
void set_error(const std::string& msg);
const std::string& get_error();
extern const char* (*g_ctx_error_hook)(const TutuCtx*);

// Validates the desc, builds the BVH if none is given, flattens to device layout.
// host_fast_tree = false leaves inner_fast empty (the caller builds the traversal tree on the device)
int flatten_scene(const TutuSceneDesc* desc, FlatScene* out, bool host_fast_tree = true);
constexpr int kFastTreeMaxDepth = 30;  // traversal trees deeper than this are rejected (stack size, trace.cuh kStackSize)
void build_host_fast_tree(FlatScene* fs);  // binned SAH over the leaves (host_scene.cpp)
// host_wide.cpp
void build_wide_tree(FlatScene* fs);
uint64_t verify_wide_tree(const FlatScene& fs);
int compute_raygen(const TutuCamera* cam, RayGen* out);
int compute_bdpt_cam(const TutuCamera* cam, BdptCamConsts* out);

}  // namespace tutu
