// ref_config_wrap.cpp — C wrapper around the REFERENCE's own util::ConfigFile, candidate::HandGeometry and
// descriptor::ImageGeometry (src/gpd/util/config_file.cpp, candidate/hand_geometry.cpp, descriptor/image_geometry.cpp: the
// source files of the reference that build without PCL / Eigen / OpenCV). Compiled together with that file FROM WHERE IT
// LIES under /root/reference into oracle/_ref/libgpd_ref_config.so (oracle/Makefile target `_ref`); nothing of the
// reference is copied into this repository. TEST INFRASTRUCTURE: pins the cfg parser of the host shim
// (gpd_b200/host) against the reference implementation (tests/test_host_cpp.py).
#include <gpd/candidate/hand_geometry.h>
#include <gpd/descriptor/image_geometry.h>
#include <gpd/util/config_file.h>

#include <cstdio>
#include <cstring>

extern "C" {

// value of `key` as the reference reads it (getValueOfKeyAsString), or `def` when the key is missing; returns 1 when the
// file was found
int gpdref_config_get(const char *file, const char *key, const char *def, char *out, int out_len) {
  gpd::util::ConfigFile cfg(file);
  const bool ok = cfg.ExtractKeys();
  const std::string v = cfg.getValueOfKeyAsString(key, def);
  std::snprintf(out, (size_t)out_len, "%s", v.c_str());
  return ok ? 1 : 0;
}
// getValueOfKey<double> / getValueOfKey<int> / getValueOfKey<bool> and the vector getters of the reference
double gpdref_config_get_double(const char *file, const char *key, double def) {
  gpd::util::ConfigFile cfg(file);
  cfg.ExtractKeys();
  return cfg.getValueOfKey<double>(key, def);
}
int gpdref_config_get_int(const char *file, const char *key, int def) {
  gpd::util::ConfigFile cfg(file);
  cfg.ExtractKeys();
  return cfg.getValueOfKey<int>(key, def);
}
int gpdref_config_get_bool(const char *file, const char *key, int def) {
  gpd::util::ConfigFile cfg(file);
  cfg.ExtractKeys();
  return cfg.getValueOfKey<bool>(key, def != 0) ? 1 : 0;
}
int gpdref_config_get_doubles(const char *file, const char *key, const char *def, double *out, int cap) {
  gpd::util::ConfigFile cfg(file);
  cfg.ExtractKeys();
  std::vector<double> v = cfg.getValueOfKeyAsStdVectorDouble(key, def);
  for (size_t i = 0; i < v.size() && (int)i < cap; i++) out[i] = v[i];
  return (int)v.size();
}
// candidate::HandGeometry(filepath) (hand_geometry.cpp:22-31): finger_width, outer_diameter, depth, height, init_bite
void gpdref_hand_geometry(const char *file, double out[5]) {
  gpd::candidate::HandGeometry g{std::string(file)};
  out[0] = g.finger_width_;
  out[1] = g.outer_diameter_;
  out[2] = g.depth_;
  out[3] = g.height_;
  out[4] = g.init_bite_;
}
// descriptor::ImageGeometry(filepath) (image_geometry.cpp:20-29): outer_diameter, depth, height | size, num_channels
void gpdref_image_geometry(const char *file, double out[3], int out2[2]) {
  gpd::descriptor::ImageGeometry g{std::string(file)};
  out[0] = g.outer_diameter_;
  out[1] = g.depth_;
  out[2] = g.height_;
  out2[0] = g.size_;
  out2[1] = g.num_channels_;
}
}
