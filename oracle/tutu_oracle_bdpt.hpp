// tutu_oracle_bdpt.hpp — CPU restatement of the reference's BDPT integrator
// (reference include/BDPT.hpp, IIntegrator.hpp:195-248, Camera.hpp:12-78, Vector.hpp:228-372).
//
// TEST INFRASTRUCTURE ONLY, #included by tutu_oracle.cpp inside its anonymous namespace.
//
// Literal port of sub_render_bdpt (BDPT.hpp:679-900, the MULTITHREAD==1 path the reference ships),
// buildEyePath (:226-293), buildLightPath (:296-390) and MISweight (:70-222), including their
// quirks (see the comments).  Random numbers: the reference's thread-local mt19937 is replaced by
// fixed slots of the Philox stream shared with the CUDA path:
//   eye vertex k (k = 1..7)   : counter depth = 32 + k, slots 0..2 -> sampleDirection draws
//   light start               : counter depth = 64,     slots 0..2 -> sampleLight, 3..4 -> sampleLightDir
//   light vertex k (k = 1..6) : counter depth = 64 + k, slots 0..2 -> sampleDirection draws

constexpr int O_MAX_PATHLENGTH = 7;  // BDPT.hpp:8
constexpr int O_RNG_EYE = 32, O_RNG_LIGHT = 64;

struct Cam {  // Camera.hpp:81-97 after initialize() (:12-48)
  V3 position, fwdDir;
  int width = 0, height = 0;
  float imagePlaneDist = 0, filmPlaneAreaInv = 0, lensAreaInv = 1;
  float w2r[16];  // world2Raster, row major

  int worldPos2PixelIndex(const V3& pos) const {  // :60-78 + raster2pxlIndex :51-58
    float r[4];
    for (int k = 0; k < 4; ++k)  // Mat4f * Vector4f, Vector.hpp:289-296
      r[k] = pos.x * w2r[4 * k + 0] + pos.y * w2r[4 * k + 1] + pos.z * w2r[4 * k + 2] + 1.f * w2r[4 * k + 3];
    float rx = r[0] / r[3], ry = r[1] / r[3];  // normalizeW
    rx -= 0.5f;
    ry -= 0.5f;
    int x = (int)rx;
    int y = (int)ry;
    if (x < 0 || x >= width || y < 0 || y >= height) return -1;
    return x + width * y;
  }
};

struct M4 {
  float e[16];
  M4() {
    for (float& v : e) v = 0;
  }
  float get(int r, int c) const { return e[c + r * 4]; }
  void set(int r, int c, float v) { e[c + r * 4] = v; }
  void setRow(int r, const V3& v, float w) { e[r * 4] = v.x, e[r * 4 + 1] = v.y, e[r * 4 + 2] = v.z, e[r * 4 + 3] = w; }
};
inline M4 mul(const M4& l, const M4& r) {  // Vector.hpp:337-349
  M4 res;
  for (int row = 0; row < 4; row++)
    for (int col = 0; col < 4; col++) {
      float acc = 0;
      for (int i = 0; i < 4; i++) acc += l.get(row, i) * r.get(i, col);
      res.set(row, col, acc);
    }
  return res;
}

Cam make_cam(const TutuCamera& c) {  // Camera.hpp:12-48
  Cam cam;
  cam.width = c.width;
  cam.height = c.height;
  cam.position = V3(c.eye[0], c.eye[1], c.eye[2]);
  V3 fwd = normalized(V3(c.viewdir[0], c.viewdir[1], c.viewdir[2]));
  V3 right = normalized(crossProduct(fwd, V3(c.updir[0], c.updir[1], c.updir[2])));
  V3 up = normalized(crossProduct(right, fwd));
  cam.fwdDir = fwd;
  V3 pos(right.dot(cam.position), up.dot(cam.position), (-fwd).dot(cam.position));
  M4 world2Cam;
  world2Cam.setRow(0, right, -pos.x);
  world2Cam.setRow(1, up, -pos.y);
  world2Cam.setRow(2, -fwd, -pos.z);
  world2Cam.setRow(3, V3(0.f), 1.f);
  // getPerspectiveMatrix(hfov, 0.1, 10000, width/height), Vector.hpp:352-372
  const float aNear = 0.1f, aFar = 10000.f, aspect = (float)c.width / c.height;
  M4 p2o;
  p2o.e[0] = aNear, p2o.e[5] = aNear, p2o.e[10] = (aNear + aFar), p2o.e[11] = aNear * aFar, p2o.e[14] = -1.0f;
  float r = tanf(((float)c.hfov_deg / 2) * 3.1415926535897f / 180) * aNear;
  float l = -r;
  float t = r / aspect;
  float b = -t;
  M4 orth_trans, orth_scale;
  orth_trans.setRow(0, V3(1, 0, 0), -(r + l) / 2);
  orth_trans.setRow(1, V3(0, 1, 0), -(t + b) / 2);
  orth_trans.setRow(2, V3(0, 0, 1), -(aNear + aFar) / 2);
  orth_trans.setRow(3, V3(0, 0, 0), 1);
  orth_scale.setRow(0, V3(2 / (r - l), 0, 0), 0);
  orth_scale.setRow(1, V3(0, 2 / -(t - b), 0), 0);
  orth_scale.setRow(2, V3(0, 0, 2 / (aNear - aFar)), 0);
  orth_scale.setRow(3, V3(0, 0, 0), 1);
  M4 perspective = mul(mul(orth_scale, orth_trans), p2o);
  M4 world2ndc = mul(perspective, world2Cam);
  M4 translate;  // getTranslate(1,1,0)
  translate.set(0, 3, 1.f), translate.set(1, 3, 1.f), translate.set(2, 3, 0.f);
  translate.set(0, 0, 1), translate.set(1, 1, 1), translate.set(2, 2, 1), translate.set(3, 3, 1);
  M4 scale;  // getScale(w/2, h/2, 0)
  scale.set(3, 3, 1), scale.set(0, 0, c.width * 0.5f), scale.set(1, 1, c.height * 0.5f), scale.set(2, 2, 0);
  M4 world2Raster = mul(scale, mul(translate, world2ndc));
  memcpy(cam.w2r, world2Raster.e, sizeof(cam.w2r));
  float tanHalfHfov = tanf((c.hfov_deg * 0.5f) * O_PI / 180.f);
  cam.imagePlaneDist = c.width / (2.f * tanHalfHfov);
  cam.filmPlaneAreaInv = 1.f / (c.width * c.height);
  cam.lensAreaInv = 1.f;
  return cam;
}

float Geo(const V3& p1, const V3& n1, const V3& p2, const V3& n2) {  // IIntegrator.hpp:223-230
  V3 p12p2 = p2 - p1;
  float dis2 = p12p2.norm2();
  p12p2 = normalized(p12p2);
  float cos = fabsf(p12p2.dot(n1));
  float cosprime = fabsf((-p12p2).dot(n2));
  return cos * cosprime / dis2;
}

float We(const V3& pos, const Cam& cam) {  // IIntegrator.hpp:233-248
  V3 inter2cam = normalized(cam.position - pos);
  int index = cam.worldPos2PixelIndex(pos);
  if (index < 0 || index >= cam.width * cam.height) return 0.f;
  float cosCamera = fabsf(cam.fwdDir.dot(-inter2cam));
  float distPixel2Cam = cam.imagePlaneDist / cosCamera;
  return distPixel2Cam * distPixel2Cam * cam.lensAreaInv * cam.filmPlaneAreaInv / (cosCamera * cosCamera);
}

bool sampleLightDir(const V3& N, float& dirPdf, V3& sampledRes, float r1, float r2) {  // IIntegrator.hpp:195-220
  float cosTheta = sqrtf(r1);
  float phi = 2 * O_PI * r2;
  V3 dir;
  float sinTheta = sqrtf(std::max(0.f, 1 - r1));
  dir.x = cosf(phi) * sinTheta;
  dir.y = sinf(phi) * sinTheta;
  dir.z = cosTheta;
  dir = normalized(dir);
  V3 res = SphereLocal2world(N, dir);
  if (normalized(res).dot(N) < 0) return false;
  dirPdf = 0.f;
  if (res.dot(N) > 0.0f) dirPdf = res.dot(N) / O_PI;
  sampledRes = res;
  return true;
}

struct PathVert {  // bdpt::eyePathVert / lightPathVert, BDPT.hpp:34-50
  V3 throughput;
  Intersection inter;
  float fwdPdf = 0, revPdf = 0, G = 0;
  bool isDelta = false;
};

struct Bdpt {
  const Scene& g;
  const Cam& cam;
  Rng rng;
  uint64_t closest_calls = 0, any_calls = 0, connections = 0;

  Intersection UpdateInter(const V3& o, const V3& d) {
    ++closest_calls;
    return getIntersection(g.root, o, d);
  }

  float MISweight(std::vector<PathVert>& epverts, std::vector<PathVert>& lpverts, int s, int t) {  // :70-222
    if (s + t == 2) return 1;
    float pdf_tEndFwd = 0, pdf_tEndRev = 0, pdf_sEndFwd = 0, pdf_sEndRev = 0, G_connect = 0;
    if (s == 0) {
      const PathVert& lightPrev = epverts[t - 2];
      const PathVert& lightvert = epverts[t - 1];
      V3 wo = normalized(lightPrev.inter.pos - lightvert.inter.pos);
      float cos = fabsf(lightvert.inter.Ng.dot(wo));
      float dirpdf = cos / O_PI;
      dirpdf = dirpdf / cos;
      float pickpdf = getLightPdf(lightvert.inter, g);
      pdf_tEndFwd = pickpdf;
      pdf_tEndRev = dirpdf;
    } else {
      const PathVert& sEndvert = lpverts[s - 1];
      const PathVert& tEndvert = epverts[t - 1];
      G_connect = Geo(sEndvert.inter.pos, sEndvert.inter.Ng, tEndvert.inter.pos, tEndvert.inter.Ng);
      const Material& sm = sEndvert.inter.mtlcolor;
      const Material& tm = tEndvert.inter.mtlcolor;
      if (t == 1) {
        V3 cam2sEnd = normalized(sEndvert.inter.pos - tEndvert.inter.pos);
        float camcos = tEndvert.inter.Ng.dot(cam2sEnd);
        float d = cam.imagePlaneDist / camcos;
        pdf_tEndFwd = (cam.filmPlaneAreaInv * d * d / camcos) / camcos;
        pdf_tEndRev = cam.lensAreaInv;
        V3 s2prev = normalized(lpverts[s - 2].inter.pos - sEndvert.inter.pos);
        pdf_sEndFwd = sm.pdf(-cam2sEnd, s2prev, sEndvert.inter.Ns, g.eta, sm.eta) / fabsf((-cam2sEnd).dot(sEndvert.inter.Ng));
        pdf_sEndRev = sm.pdf(s2prev, -cam2sEnd, sEndvert.inter.Ns, g.eta, sm.eta) / fabsf(s2prev.dot(sEndvert.inter.Ng));
      } else if (s == 1) {
        V3 light2tEnd = normalized(tEndvert.inter.pos - sEndvert.inter.pos);
        float cos = sEndvert.inter.Ng.dot(light2tEnd);
        pdf_sEndFwd = cos / O_PI / cos;
        pdf_sEndRev = sEndvert.revPdf;
        V3 t2prev = normalized(epverts[t - 2].inter.pos - tEndvert.inter.pos);
        pdf_tEndFwd = tm.pdf(-light2tEnd, t2prev, tEndvert.inter.Ns, g.eta, tm.eta) / fabsf((-light2tEnd).dot(tEndvert.inter.Ng));
        pdf_tEndRev = tm.pdf(t2prev, -light2tEnd, tEndvert.inter.Ns, g.eta, tm.eta) / fabsf(t2prev.dot(tEndvert.inter.Ng));
      } else {
        V3 s2t = normalized(tEndvert.inter.pos - sEndvert.inter.pos);
        V3 s2prev = normalized(lpverts[s - 2].inter.pos - sEndvert.inter.pos);
        V3 t2prev = normalized(epverts[t - 2].inter.pos - tEndvert.inter.pos);
        pdf_sEndFwd = sm.pdf(s2t, s2prev, sEndvert.inter.Ns, g.eta, sm.eta) / fabsf(s2t.dot(sEndvert.inter.Ng));
        pdf_sEndRev = sm.pdf(s2prev, s2t, sEndvert.inter.Ns, g.eta, sm.eta) / fabsf(s2prev.dot(sEndvert.inter.Ng));
        pdf_tEndFwd = tm.pdf(-s2t, t2prev, tEndvert.inter.Ns, g.eta, tm.eta) / fabsf((-s2t).dot(tEndvert.inter.Ng));
        pdf_tEndRev = tm.pdf(t2prev, -s2t, tEndvert.inter.Ns, g.eta, tm.eta) / fabsf(t2prev.dot(tEndvert.inter.Ng));
      }
    }
    struct Node {
      float toLight = 0, toEye = 0;
      bool isDelta = false;
    };
    Node mis[2 * O_MAX_PATHLENGTH + 4];
    int k = s + t - 1;
    for (int i = 0; i < s - 1; ++i) {
      mis[i].toLight = (i == 0) ? lpverts[0].revPdf : lpverts[i].revPdf * lpverts[i].G;
      mis[i].toEye = lpverts[i].fwdPdf * lpverts[i + 1].G;
      mis[i].isDelta = lpverts[i].isDelta;
    }
    if (s > 0) {
      mis[s - 1].toLight = (s == 1) ? pdf_sEndRev : pdf_sEndRev * lpverts[s - 1].G;
      mis[s - 1].toEye = pdf_sEndFwd * G_connect;
      mis[s - 1].isDelta = lpverts[s - 1].isDelta;
    }
    for (int ti = 0; ti < t - 1; ++ti) {
      mis[k - ti].toEye = (ti == 0) ? epverts[ti].revPdf : epverts[ti].revPdf * epverts[ti].G;
      mis[k - ti].toLight = epverts[ti].fwdPdf * epverts[ti + 1].G;
      mis[k - ti].isDelta = epverts[ti].isDelta;
    }
    mis[k - (t - 1)].toEye = (t == 1) ? pdf_tEndRev : pdf_tEndRev * epverts[t - 1].G;
    mis[k - (t - 1)].toLight = (s == 0) ? pdf_tEndFwd : pdf_tEndFwd * G_connect;
    mis[k - (t - 1)].isDelta = epverts[t - 1].isDelta;

    float p_i_plus_1 = 1.0f;
    float denominator = 1.0f;
    for (int i = s; i < k; ++i) {
      if (i == 0) {
        p_i_plus_1 *= mis[0].toLight / mis[1].toLight;
        if (mis[1].isDelta) continue;
      } else {
        p_i_plus_1 *= mis[i - 1].toEye / mis[i + 1].toLight;
        if (mis[i].isDelta || mis[i + 1].isDelta) continue;
      }
      denominator += p_i_plus_1 * p_i_plus_1;
    }
    float p_i_minus_1 = 1.0f;
    for (int i = s; i > 0; --i) {
      if (i == (k + 1)) {
      } else if (i == 1) {
        p_i_minus_1 *= mis[1].toLight / mis[0].toLight;
        if (mis[0].isDelta) continue;
      } else {
        p_i_minus_1 *= mis[i].toLight / mis[i - 2].toEye;
        if (mis[i - 1].isDelta || mis[i - 2].isDelta) continue;
      }
      denominator += p_i_minus_1 * p_i_minus_1;
    }
    float res = 1 / denominator;
    if (res < O_MIN_DIVISOR || std::isnan(res) || std::isinf(res)) return 0;
    return 1 / denominator;
  }

  // one step of either random walk (the loop bodies of buildEyePath :236-292 and buildLightPath
  // :334-389 are the same code up to the adjoint flag).  Returns false when the walk ends.
  bool walk_step(std::vector<PathVert>& verts, Intersection& nxtInter, V3& tp, V3& wi, bool adjoint, int rng_depth) {
    PathVert v;
    v.inter = nxtInter;
    v.throughput = tp;
    if (v.inter.obj->isTextureActivated) textureModify(v.inter, g);
    V3 wo = -wi;
    auto [success, TIR] = v.inter.mtlcolor.sampleDirection(wo, v.inter.Ns, wi, g.eta,
                                                           [&](int k) { return rng.get(rng_depth, k); });
    if (!success) return false;
    wi = normalized(wi);
    float dirPdf = v.inter.mtlcolor.pdf(wi, wo, v.inter.Ns, g.eta, v.inter.mtlcolor.eta);
    if (TIR) {
      wi = normalized(getReflectionDir(wo, v.inter.Ns));
      dirPdf = 1;
    }
    if (dirPdf == 0) return false;
    float cos = fabsf(wi.dot(v.inter.Ng));
    v.fwdPdf = dirPdf / cos;
    if (v.inter.mtlcolor.mType == TUTU_MAT_PERFECT_REFLECTIVE || v.inter.mtlcolor.mType == TUTU_MAT_PERFECT_REFRACTIVE) {
      v.revPdf = v.fwdPdf;
      v.isDelta = true;
    } else {
      v.revPdf = v.inter.mtlcolor.pdf(wo, wi, v.inter.Ns, g.eta, v.inter.mtlcolor.eta);
      v.revPdf = v.revPdf / fabsf(wo.dot(v.inter.Ng));
      v.isDelta = false;
    }
    const PathVert& pre = verts.back();
    v.G = Geo(pre.inter.pos, pre.inter.Ng, v.inter.pos, v.inter.Ng);
    verts.emplace_back(v);
    if (v.inter.mtlcolor.hasEmission()) return false;
    V3 bsdf = v.inter.mtlcolor.BxDF(wi, wo, v.inter.Ng, v.inter.Ns, g.eta, adjoint, TIR);
    if (dirPdf < O_MIN_DIVISOR) return false;
    tp = tp * bsdf * cos / dirPdf;
    V3 orig = v.inter.pos;
    bool rayInside = v.inter.Ns.dot(wi) < 0;
    offsetRayOrig(orig, v.inter.Ns, rayInside);
    nxtInter = UpdateInter(orig, wi);
    return nxtInter.intersected;
  }

  void buildEyePath(std::vector<PathVert>& epverts) {  // :226-293
    V3 tp = epverts[1].throughput;
    Intersection nxtInter = epverts[1].inter;
    V3 wi = normalized(epverts[1].inter.pos - epverts[0].inter.pos);
    epverts.pop_back();
    while ((int)epverts.size() < O_MAX_PATHLENGTH + 1) {
      const int k = (int)epverts.size();
      if (!walk_step(epverts, nxtInter, tp, wi, false, O_RNG_EYE + k)) return;
    }
  }

  void buildLightPath(std::vector<PathVert>& lpverts) {  // :296-390
    Intersection lightInter;
    float pickpdf;
    sampleLight(lightInter, pickpdf, g, rng, O_RNG_LIGHT);
    V3 tp = V3(1 / pickpdf);
    PathVert lpv;
    lpv.inter = lightInter;
    lpv.throughput = tp;
    lpv.revPdf = pickpdf;
    lpv.isDelta = false;
    float dirPdf;
    V3 wi;
    if (!sampleLightDir(lightInter.Ng, dirPdf, wi, rng.get(O_RNG_LIGHT, 3), rng.get(O_RNG_LIGHT, 4))) return;
    wi = normalized(wi);
    float wi_n_cos = fabsf(wi.dot(lightInter.Ng));
    lpv.fwdPdf = dirPdf / wi_n_cos;
    lpverts.emplace_back(lpv);
    tp = lpverts[0].throughput * wi_n_cos / dirPdf;
    V3 orig = lightInter.pos;
    offsetRayOrig(orig, lightInter.Ns, false);
    Intersection nxtInter = UpdateInter(orig, wi);
    if (!nxtInter.intersected) return;
    if (nxtInter.mtlcolor.hasEmission()) return;
    while ((int)lpverts.size() < O_MAX_PATHLENGTH) {
      const int k = (int)lpverts.size();
      if (!walk_step(lpverts, nxtInter, tp, wi, true, O_RNG_LIGHT + k)) return;
    }
  }

  bool shadowBlocked(V3 orig, const V3& target) {
    ++any_calls;
    return isShadowRayBlocked(orig, target, g);
  }

  // One sample of sub_render_bdpt's inner loop (:707-888).  `estimate` collects the s != .. / t >= 2
  // strategies of this pixel; t == 1 strategies are splatted through `splat(index, value)`.
  // Returns false when the reference `break`s out of the sample loop (primary ray missed, :733-734).
  template <class Splat>
  bool sample(const V3& eyePos, const V3& pixelPos, const V3& rayDir, float SPP_inv, V3& estimate, Splat&& splat) {
    std::vector<PathVert> epverts, lpverts;
    V3 wi = rayDir;
    PathVert ev;
    ev.inter.pos = eyePos;
    ev.inter.intersected = true;
    ev.inter.Ng = cam.fwdDir;
    ev.throughput = V3(1.f);
    ev.revPdf = cam.lensAreaInv;
    float wi_n_cos = fabsf(wi.dot(cam.fwdDir));
    float d2 = (pixelPos - cam.position).norm2();
    ev.fwdPdf = d2 * cam.filmPlaneAreaInv / wi_n_cos;
    ev.fwdPdf = ev.fwdPdf / wi_n_cos;
    ev.isDelta = false;
    epverts.emplace_back(ev);
    float pdfCam_w = d2 * cam.lensAreaInv * cam.filmPlaneAreaInv / wi_n_cos;
    V3 tp = epverts[0].throughput * wi_n_cos / pdfCam_w;
    Intersection eVert2 = UpdateInter(eyePos, wi);
    if (!eVert2.intersected) return false;
    ev.inter = eVert2;
    ev.throughput = tp;
    epverts.emplace_back(ev);
    buildEyePath(epverts);
    buildLightPath(lpverts);
    float we = We(pixelPos, cam);
    V3 contrib;
    if (epverts.size() < 2) return true;
    for (int pathLength = 1; pathLength <= O_MAX_PATHLENGTH; pathLength++) {
      for (int s = 0; s < pathLength + 1; s++) {
        int t = pathLength + 1 - s;
        if (t <= 0 || t > (int)epverts.size() || s > (int)lpverts.size()) continue;
        if (s == 0) {
          // (:767) tests the FIRST hit's material (`ev` of the enclosing scope); an UNLIT first hit
          // never gets here because its sampleDirection fails and epverts.size() stays 1
          if (ev.inter.mtlcolor.mType == TUTU_MAT_UNLIT) {
            estimate = estimate + ev.inter.mtlcolor.diffuse;
            continue;
          }
          const PathVert& e = epverts[t - 1];
          if (!e.inter.mtlcolor.hasEmission()) continue;
          V3 l = e.inter.mtlcolor.emission;
          contrib = we * e.throughput * l;
          if (contrib.norm2() == 0) continue;
          if (std::isnan(contrib.x)) continue;
          float misw = MISweight(epverts, lpverts, s, t);
          estimate = estimate + misw * contrib;
          continue;
        }
        if (t == 1) {
          const PathVert& lv = lpverts[s - 1];
          if (lv.inter.mtlcolor.hasEmission()) continue;
          V3 l = lpverts[0].inter.mtlcolor.emission;
          V3 orig = lv.inter.pos;
          V3 wi2 = normalized(cam.position - orig);
          V3 wo;
          bool rayInside;
          V3 bsdf;
          if (s == 1) {
            bsdf = V3(1);
            rayInside = false;
          } else {
            wo = normalized(lpverts[s - 2].inter.pos - lv.inter.pos);
            rayInside = wi2.dot(lv.inter.Ng) < 0;  // (:803) Ng here, Ns in the single-thread twin (:528)
            bsdf = lv.inter.mtlcolor.BxDF(wi2, wo, lv.inter.Ng, lv.inter.Ns, g.eta, true);
          }
          V3 camPos = cam.position;
          float G = Geo(camPos, cam.fwdDir, lv.inter.pos, lv.inter.Ng);
          float we2 = We(lv.inter.pos, cam);
          contrib = l * bsdf * lv.throughput * G * we2 * SPP_inv;
          if (contrib.norm2() == 0) continue;
          if (std::isnan(contrib.x)) continue;
          float misw = MISweight(epverts, lpverts, s, t);
          offsetRayOrig(orig, lv.inter.Ns, rayInside);
          if (!shadowBlocked(orig, cam.position) && wi2.dot(cam.fwdDir) < 0) {
            int index = cam.worldPos2PixelIndex(lv.inter.pos);
            splat(index, misw * contrib);
          }
          continue;
        }
        const PathVert& lv = lpverts[s - 1];
        V3 l = lpverts[0].inter.mtlcolor.emission;
        const PathVert& e = epverts[t - 1];
        if (e.inter.mtlcolor.hasEmission()) continue;
        V3 connectDir = normalized(e.inter.pos - lv.inter.pos);
        V3 e_wo = normalized(epverts[t - 2].inter.pos - e.inter.pos);
        V3 evBSDF = e.inter.mtlcolor.BxDF(-connectDir, e_wo, e.inter.Ng, e.inter.Ns, g.eta, false);
        V3 lvBSDF;
        V3 l_wo;
        if (s == 1) {
          if (connectDir.dot(lv.inter.Ns) >= 0) lvBSDF = V3(1.f);
          else
            lvBSDF = V3(0);
        } else {
          l_wo = normalized(lpverts[s - 2].inter.pos - lv.inter.pos);
          lvBSDF = lv.inter.mtlcolor.BxDF(connectDir, l_wo, lv.inter.Ng, lv.inter.Ns, g.eta, true);
        }
        V3 eOrig = e.inter.pos;
        bool rayInside = e_wo.dot(e.inter.Ns) < 0;
        offsetRayOrig(eOrig, e.inter.Ns, rayInside);
        V3 lorig = lv.inter.pos;
        if (s == 1) {
          offsetRayOrig(lorig, lv.inter.Ns, false);
        } else {
          rayInside = l_wo.dot(lv.inter.Ns) < 0;
          offsetRayOrig(lorig, lv.inter.Ns, rayInside);
        }
        ++connections;
        if (shadowBlocked(eOrig, lorig)) continue;
        float G = Geo(e.inter.pos, e.inter.Ng, lv.inter.pos, lv.inter.Ng);
        contrib = we * e.throughput * evBSDF * G * lv.throughput * lvBSDF * l;
        if (contrib.norm2() == 0) continue;
        if (std::isnan(contrib.x)) continue;
        float misw = MISweight(epverts, lpverts, s, t);
        estimate = estimate + misw * contrib;
      }
    }
    return true;
  }
};
