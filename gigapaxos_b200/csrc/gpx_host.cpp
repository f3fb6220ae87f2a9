/*
 * gpx_host.cpp -- host-only pieces of the C ABI: the gigapaxos.properties reader.
 *
 * Mirrors the reference's config surface (utils/Config.java:231-347,
 * gigapaxos/PaxosConfig.java:64,83-90,156-170): a java.util.Properties file whose keys are
 * the PC enum names; unknown keys are ignored, missing keys keep the PC defaults.  Node map
 * entries `active.<name>=host:port` are returned through gpx_properties_actives.
 */
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "gpx.h"

extern "C" void gpx_config_defaults(gpx_config* c);

namespace {

std::string trim(const std::string& s) {
  size_t a = 0, b = s.size();
  while (a < b && isspace((unsigned char)s[a])) a++;
  while (b > a && isspace((unsigned char)s[b - 1])) b--;
  return s.substr(a, b - a);
}

/* java.util.Properties.load subset: '#'/'!' comments, key[=:]value, trailing '\' continuation */
bool load_properties(const char* path, std::map<std::string, std::string>& out) {
  FILE* f = fopen(path, "r");
  if (!f) return false;
  std::string line, logical;
  char buf[4096];
  std::vector<std::string> lines;
  while (fgets(buf, sizeof buf, f)) {
    line = buf;
    while (!line.empty() && (line.back() == '\n' || line.back() == '\r')) line.pop_back();
    std::string t = trim(line);
    if (logical.empty() && (t.empty() || t[0] == '#' || t[0] == '!')) continue;
    if (!t.empty() && t.back() == '\\') {
      t.pop_back();
      logical += t;
      continue;
    }
    logical += t;
    lines.push_back(logical);
    logical.clear();
  }
  if (!logical.empty()) lines.push_back(logical);
  fclose(f);
  for (const std::string& l : lines) {
    size_t p = l.find_first_of("=:");
    std::string k, v;
    if (p == std::string::npos) {
      k = trim(l);
    } else {
      k = trim(l.substr(0, p));
      v = trim(l.substr(p + 1));
    }
    if (!k.empty()) out[k] = v;
  }
  return true;
}

bool as_bool(const std::string& v) {
  std::string t;
  for (char c : v) t += (char)tolower((unsigned char)c);
  return t == "true" || t == "1" || t == "yes";
}

std::map<std::string, std::string> g_last_actives;

}  // namespace

extern "C" {

int gpx_config_from_properties(const char* path, gpx_config* cfg) {
  if (!path || !cfg) return GPX_EINVAL;
  gpx_config_defaults(cfg);
  std::map<std::string, std::string> p;
  if (!load_properties(path, p)) return GPX_EIO;
  auto has = [&](const char* k) { return p.find(k) != p.end(); };
  if (has("BATCHING_ENABLED")) cfg->batching_enabled = as_bool(p["BATCHING_ENABLED"]);
  if (has("MAX_BATCH_SIZE")) cfg->max_batch_size = atoi(p["MAX_BATCH_SIZE"].c_str());
  if (has("CHECKPOINT_INTERVAL")) cfg->checkpoint_interval = atoi(p["CHECKPOINT_INTERVAL"].c_str());
  if (has("CPI_NOISE")) cfg->cpi_noise = atof(p["CPI_NOISE"].c_str());
  if (has("GC_MAJORITY_EXECUTED")) cfg->gc_majority_executed = as_bool(p["GC_MAJORITY_EXECUTED"]);
  if (has("LOG_META_DECISIONS")) cfg->log_meta_decisions = as_bool(p["LOG_META_DECISIONS"]);
  if (has("ENABLE_JOURNALING")) cfg->journaling_enabled = as_bool(p["ENABLE_JOURNALING"]);
  if (has("DISABLE_LOGGING") && as_bool(p["DISABLE_LOGGING"]) && has("ENABLE_JOURNALING") &&
      !as_bool(p["ENABLE_JOURNALING"]))
    cfg->journaling_enabled = 0; /* GET_ACCEPTED_PVALUES_FROM_DISK = logging || journaling */
  if (has("PINSTANCES_CAPACITY")) cfg->max_groups = (uint32_t)strtoul(p["PINSTANCES_CAPACITY"].c_str(), nullptr, 10);
  if (has("MAX_GROUP_SIZE")) {
    int v = atoi(p["MAX_GROUP_SIZE"].c_str());
    if (v > 0 && v <= GPX_MAX_GROUP_SIZE) cfg->max_group_size = (uint32_t)v;
  }
  long long max_payload = 4ll * 1024 * 1024, max_log = 5ll * 1024 * 1024;
  if (has("NIO_MAX_PAYLOAD_SIZE")) max_payload = atoll(p["NIO_MAX_PAYLOAD_SIZE"].c_str());
  if (has("MAX_LOG_MESSAGE_SIZE")) max_log = atoll(p["MAX_LOG_MESSAGE_SIZE"].c_str());
  cfg->max_batch_bytes = max_payload < max_log ? max_payload : max_log; /* RequestBatcher.java:204-208 */
  g_last_actives.clear();
  for (auto& kv : p)
    if (kv.first.rfind("active.", 0) == 0) g_last_actives[kv.first.substr(7)] = kv.second;
  return GPX_OK;
}

/* `active.<name>=host:port` entries of the last parsed file, as "name=host:port\n" lines */
int gpx_properties_actives(char* out, size_t cap) {
  std::string s;
  for (auto& kv : g_last_actives) s += kv.first + "=" + kv.second + "\n";
  if (s.size() + 1 > cap) return GPX_ERANGE;
  memcpy(out, s.c_str(), s.size() + 1);
  return GPX_OK;
}

} /* extern "C" */
