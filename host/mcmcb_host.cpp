// Host-side mirror of the reference's driver pieces (see mcmcb_host.h).  Pure C++17, no CUDA.
#include "mcmcb_host.h"

#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

namespace {
thread_local std::string g_err;

int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}

std::string lower(std::string s) {
  std::transform(s.begin(), s.end(), s.begin(), [](unsigned char c) { return (char)std::tolower(c); });
  return s;
}
std::string trim(const std::string& s) {
  size_t a = 0, b = s.size();
  while (a < b && std::isspace((unsigned char)s[a])) a++;
  while (b > a && std::isspace((unsigned char)s[b - 1])) b--;
  return s.substr(a, b - a);
}

// One namelist group as key -> raw value text.  Handles `!` comments (outside quotes), quoted strings,
// several assignments per line separated by commas, the terminating `/`.
int read_group(const std::string& text, const std::string& group, std::map<std::string, std::string>& kv, bool* found) {
  *found = false;
  const std::string ltext = lower(text);
  size_t pos = 0;
  const std::string tag = "&" + lower(group);
  for (;;) {
    pos = ltext.find(tag, pos);
    if (pos == std::string::npos) return 0;
    const size_t after = pos + tag.size();
    const bool line_start = (pos == 0) || ltext.find_last_of('\n', pos) == std::string::npos ||
                            trim(ltext.substr(ltext.find_last_of('\n', pos) + 1, pos - ltext.find_last_of('\n', pos) - 1)).empty();
    if (line_start && (after >= ltext.size() || !(std::isalnum((unsigned char)ltext[after]) || ltext[after] == '_'))) break;
    pos = after;
  }
  *found = true;
  // strip comments and find the end of the group
  std::string body;
  bool inq = false;
  char q = 0;
  bool incomment = false, closed = false;
  for (size_t i = pos + tag.size(); i < text.size(); i++) {
    const char c = text[i];
    if (incomment) {
      if (c == '\n') { incomment = false; body += '\n'; }
      continue;
    }
    if (inq) {
      body += c;
      if (c == q) inq = false;
      continue;
    }
    if (c == '\'' || c == '"') { inq = true; q = c; body += c; continue; }
    if (c == '!') { incomment = true; continue; }
    if (c == '/') { closed = true; break; }
    body += c;
  }
  if (!closed) return fail(MCMCBH_EPARSE, "namelist group &" + group + " is not terminated by /");
  // split into key = value: a key is an identifier followed by '='
  size_t i = 0;
  while (i < body.size()) {
    while (i < body.size() && (std::isspace((unsigned char)body[i]) || body[i] == ',')) i++;
    if (i >= body.size()) break;
    size_t k0 = i;
    while (i < body.size() && (std::isalnum((unsigned char)body[i]) || body[i] == '_')) i++;
    const std::string key = lower(body.substr(k0, i - k0));
    while (i < body.size() && std::isspace((unsigned char)body[i])) i++;
    if (key.empty() || i >= body.size() || body[i] != '=')
      return fail(MCMCBH_EPARSE, "namelist &" + group + ": expected `name = value` near `" + body.substr(k0, 20) + "`");
    i++;
    // value runs to the next `identifier =` or the end; quotes protect everything
    std::string val;
    bool vq = false;
    char vqc = 0;
    while (i < body.size()) {
      const char c = body[i];
      if (vq) { val += c; if (c == vqc) vq = false; i++; continue; }
      if (c == '\'' || c == '"') { vq = true; vqc = c; val += c; i++; continue; }
      if (std::isalpha((unsigned char)c) || c == '_') {  // maybe the next key (values like .true. / 1e5 start otherwise)
        size_t j = i;
        while (j < body.size() && (std::isalnum((unsigned char)body[j]) || body[j] == '_')) j++;
        size_t j2 = j;
        while (j2 < body.size() && std::isspace((unsigned char)body[j2])) j2++;
        const bool prev_sep = val.empty() || std::isspace((unsigned char)val.back()) || val.back() == ',';
        if (j2 < body.size() && body[j2] == '=' && prev_sep) break;
      }
      val += c;
      i++;
    }
    val = trim(val);
    while (!val.empty() && val.back() == ',') val = trim(val.substr(0, val.size() - 1));
    kv[key] = val;
  }
  return 0;
}

bool parse_int(const std::string& v, int& out) {
  char* e = nullptr;
  const double d = std::strtod(v.c_str(), &e);  // Fortran list-directed read of `0` into an integer; be lenient on `1.`
  if (e == v.c_str() || *trim(std::string(e)).c_str() != 0) return false;
  out = (int)d;
  return (double)out == d;
}
bool parse_ll(const std::string& v, long long& out) {
  char* e = nullptr;
  out = std::strtoll(v.c_str(), &e, 10);
  return e != v.c_str() && trim(std::string(e)).empty();
}
bool parse_double(std::string v, double& out) {
  for (auto& c : v)
    if (c == 'd' || c == 'D') c = 'e';  // Fortran double-precision exponent
  const size_t us = v.find('_');        // kind suffix, e.g. 2.5_dbl
  if (us != std::string::npos) v = v.substr(0, us);
  char* e = nullptr;
  out = std::strtod(v.c_str(), &e);
  return e != v.c_str() && trim(std::string(e)).empty();
}
bool parse_string(const std::string& v, char* out, size_t cap) {
  std::string s = v;
  if (s.size() >= 2 && (s.front() == '\'' || s.front() == '"') && s.back() == s.front()) s = s.substr(1, s.size() - 2);
  s = trim(s);
  if (s.size() + 1 > cap) return false;
  std::memset(out, 0, cap);
  std::memcpy(out, s.c_str(), s.size());
  return true;
}

void set_str(char* dst, const char* s) {
  std::memset(dst, 0, MCMCBH_PATH);
  std::strncpy(dst, s, MCMCBH_PATH - 1);
}

std::string join(const char* dir, const char* f) {
  if (!dir || !*dir || f[0] == '/') return f;
  return std::string(dir) + "/" + f;
}
}  // namespace

extern "C" const char* mcmcbh_last_error(void) { return g_err.c_str(); }
extern "C" void mcmcbh_free(void* p) { std::free(p); }

extern "C" int mcmcbh_read_namelist(const char* path, mcmcb_config* cfg, mcmcbh_files* f) {
  if (!path || !cfg || !f) return fail(MCMCBH_EINVAL, "null argument");
  // ---- defaults, MCMC_init_namelist (mcmcinit.F90:184-230).  The kernel-relevant ones live in the C ABI's
  // own default routine in the CUDA library; they are restated here so that this file needs no GPU library.
  std::memset(cfg, 0, sizeof *cfg);
  cfg->abi_version = MCMCB_ABI_VERSION;
  cfg->method = MCMCB_DRAM; cfg->nsimu = 0; cfg->doadapt = 1; cfg->doburnin = 0; cfg->burnintime = 0;
  cfg->badaptint = -1; cfg->greedy = 0; cfg->scalelimit = 0.05; cfg->scalefactor = 2.5; cfg->drscale = 0.0;
  cfg->adaptint = 100; cfg->adapthist = 0; cfg->adaptend = 0; cfg->initcmatn = 0; cfg->N0 = 1.0; cfg->S02 = 0.0;
  cfg->updatesigma = 1; cfg->condmax = 0.0; cfg->alphatarget = 0.234; cfg->nuparam = 0.7;
  cfg->nchains = 1; cfg->chain_offset = 0; cfg->seed = 0; cfg->rng_mode = MCMCB_RNG_PHILOX; cfg->device = 0;
  cfg->store_chains = 1; cfg->lanes_per_chain = 0; cfg->dump_stride = 0; cfg->kernel = 0; cfg->pool_adapt = 0;
  cfg->diag_stride = 0; cfg->diag_lags = 8; cfg->ngpus = 1;
  std::strcpy(cfg->model, "expreg");
  std::memset(f, 0, sizeof *f);
  set_str(f->chainfile, "chain.dat"); set_str(f->s2file, "s2chain.dat"); set_str(f->ssfile, "sschain.dat");
  set_str(f->priorsfile, ""); set_str(f->cov0file, "mcmccov.dat"); set_str(f->covffile, "mcmccovf.dat");
  set_str(f->covnfile, ""); set_str(f->meanfile, "mcmcmean.dat"); set_str(f->nmlffile, "");
  set_str(f->parfile, "mcmcpar.dat"); set_str(f->parffile, "mcmcparf.dat"); set_str(f->sigma2file, "mcmcsigma2.dat");
  set_str(f->sigma2ffile, "mcmcsigma2f.dat"); set_str(f->datafile, "data.dat");
  f->verbosity = 1; f->printint = 500; f->dumpint = 0; f->usrfunlen = 0; f->filepars = 1;
  f->svddim = 0; f->sstype = 0; f->condmaxini = 1.0e15; f->sstrans = -1.0;

  std::ifstream in(path);
  if (!in) return fail(MCMCBH_ENOFILE, std::string("File ") + path + " not found, no MCMC run");  // mcmcinit.F90:121-125
  std::stringstream ss;
  ss << in.rdbuf();
  const std::string text = ss.str();

  std::map<std::string, std::string> kv;
  bool found = false;
  int rc = read_group(text, "mcmc", kv, &found);
  if (rc) return rc;
  if (!found) return fail(MCMCBH_EPARSE, "namelist group &mcmc not found");
  int sstype = 0;          // mcmcinit.F90:222-223: legacy likelihood selector of the Modest interface
  double sstrans = -1.0;
  for (auto& [k, v] : kv) {
    bool ok = true;
    if (k == "nsimu") ok = parse_int(v, cfg->nsimu);
    else if (k == "doadapt") ok = parse_int(v, cfg->doadapt);
    else if (k == "doburnin") ok = parse_int(v, cfg->doburnin);
    else if (k == "adaptint") ok = parse_int(v, cfg->adaptint);
    else if (k == "adapthist") ok = parse_int(v, cfg->adapthist);
    else if (k == "scalelimit") ok = parse_double(v, cfg->scalelimit);
    else if (k == "scalefactor") ok = parse_double(v, cfg->scalefactor);
    else if (k == "drscale") ok = parse_double(v, cfg->drscale);
    else if (k == "badaptint") ok = parse_int(v, cfg->badaptint);
    else if (k == "adaptend") ok = parse_int(v, cfg->adaptend);
    else if (k == "initcmatn") ok = parse_int(v, cfg->initcmatn);
    else if (k == "n0") ok = parse_double(v, cfg->N0);
    else if (k == "s02") ok = parse_double(v, cfg->S02);
    else if (k == "filepars") ok = parse_int(v, f->filepars);
    else if (k == "burnintime") ok = parse_int(v, cfg->burnintime);
    else if (k == "greedy") ok = parse_int(v, cfg->greedy);
    else if (k == "printint") ok = parse_int(v, f->printint);
    else if (k == "updatesigma") ok = parse_int(v, cfg->updatesigma);
    else if (k == "usrfunlen") ok = parse_int(v, f->usrfunlen);
    else if (k == "chainfile") ok = parse_string(v, f->chainfile, MCMCBH_PATH);
    else if (k == "s2file") ok = parse_string(v, f->s2file, MCMCBH_PATH);
    else if (k == "ssfile") ok = parse_string(v, f->ssfile, MCMCBH_PATH);
    else if (k == "svddim") ok = parse_int(v, f->svddim);
    else if (k == "condmax") ok = parse_double(v, cfg->condmax);
    else if (k == "cov0file") ok = parse_string(v, f->cov0file, MCMCBH_PATH);
    else if (k == "covffile") ok = parse_string(v, f->covffile, MCMCBH_PATH);
    else if (k == "covnfile") ok = parse_string(v, f->covnfile, MCMCBH_PATH);
    else if (k == "meanfile") ok = parse_string(v, f->meanfile, MCMCBH_PATH);
    else if (k == "nmlffile") ok = parse_string(v, f->nmlffile, MCMCBH_PATH);
    else if (k == "parfile") ok = parse_string(v, f->parfile, MCMCBH_PATH);
    else if (k == "parffile") ok = parse_string(v, f->parffile, MCMCBH_PATH);
    else if (k == "sigma2file") ok = parse_string(v, f->sigma2file, MCMCBH_PATH);
    else if (k == "sigma2ffile") ok = parse_string(v, f->sigma2ffile, MCMCBH_PATH);
    else if (k == "condmaxini") ok = parse_double(v, f->condmaxini);
    else if (k == "sstrans") { ok = parse_double(v, sstrans); f->sstrans = sstrans; }
    else if (k == "sstype") { ok = parse_int(v, sstype); f->sstype = sstype; }
    else if (k == "dumpint") ok = parse_int(v, f->dumpint);
    else if (k == "priorsfile") ok = parse_string(v, f->priorsfile, MCMCBH_PATH);
    else if (k == "verbosity") ok = parse_int(v, f->verbosity);
    else if (k == "method") {
      char m[16];
      ok = parse_string(v, m, sizeof m);
      const std::string ms = lower(m);
      if (ms == "dram" || ms == "am") cfg->method = MCMCB_DRAM;  // anything that is not scam/ram/er runs MCMC_run (mcmc_main.F90:29-37)
      else if (ms == "ram") cfg->method = MCMCB_RAM;
      else if (ms == "scam") cfg->method = MCMCB_SCAM;
      else if (ms == "er") cfg->method = MCMCB_ER;  // MCMC_run_er (mcmc_main.F90:31-32)
      else cfg->method = MCMCB_DRAM;
    }
    else if (k == "alphatarget") ok = parse_double(v, cfg->alphatarget);
    else if (k == "nuparam") ok = parse_double(v, cfg->nuparam);
    else return fail(MCMCBH_EPARSE, "namelist &mcmc: unknown variable `" + k + "`");
    if (!ok) return fail(MCMCBH_EPARSE, "namelist &mcmc: bad value for `" + k + "`: " + v);
  }
  // check_mcmcinit_parameters' side effects of sstype on the sampler (mcmcinit.F90:271-321); the ss types themselves
  // belong to the Modest interface (`mdstlsqs`), not to the standalone library, and have no device model here
  if (sstype == 3) { cfg->updatesigma = 0; cfg->S02 = 1.0; }               // Poisson likelihood
  else if (sstype == 5) {                                                   // t with fixed df = sstrans
    if (sstrans < 1.0) return fail(MCMCBH_EPARSE, "ERROR: please set sstrans = df, when sstype = 5");
    cfg->updatesigma = 0;
  } else if (sstype == 6) { cfg->updatesigma = 0; cfg->S02 = std::sqrt(cfg->S02); }  // Laplace
  kv.clear();
  rc = read_group(text, "mcmcb", kv, &found);
  if (rc) return rc;
  for (auto& [k, v] : kv) {
    bool ok = true;
    long long ll = 0;
    if (k == "nchains") { ok = parse_ll(v, ll); cfg->nchains = ll; }
    else if (k == "chain_offset") { ok = parse_ll(v, ll); cfg->chain_offset = ll; }
    else if (k == "seed") { ok = parse_ll(v, ll); cfg->seed = (unsigned long long)ll; }
    else if (k == "device") ok = parse_int(v, cfg->device);
    else if (k == "store_chains") ok = parse_int(v, cfg->store_chains);
    else if (k == "lanes_per_chain") ok = parse_int(v, cfg->lanes_per_chain);
    else if (k == "dump_stride") ok = parse_int(v, cfg->dump_stride);
    else if (k == "kernel") ok = parse_int(v, cfg->kernel);
    else if (k == "pool_adapt") ok = parse_int(v, cfg->pool_adapt);
    else if (k == "diag_stride") ok = parse_int(v, cfg->diag_stride);
    else if (k == "diag_lags") ok = parse_int(v, cfg->diag_lags);
    else if (k == "ngpus") ok = parse_int(v, cfg->ngpus);
    else if (k == "model") ok = parse_string(v, cfg->model, sizeof cfg->model);
    else if (k == "datafile") ok = parse_string(v, f->datafile, MCMCBH_PATH);
    else return fail(MCMCBH_EPARSE, "namelist &mcmcb: unknown variable `" + k + "`");
    if (!ok) return fail(MCMCBH_EPARSE, "namelist &mcmcb: bad value for `" + k + "`: " + v);
  }
  return MCMCBH_OK;
}

extern "C" int mcmcbh_load_dat(const char* path, double** data, int* rows, int* cols) {
  if (!path || !data || !rows || !cols) return fail(MCMCBH_EINVAL, "null argument");
  std::ifstream in(path);
  if (!in) return fail(MCMCBH_ENOFILE, std::string("cannot open ") + path);
  std::vector<double> v;
  int nr = 0, nc = -1;
  std::string line;
  while (std::getline(in, line)) {
    std::string s = trim(line);
    if (s.empty()) continue;
    if (std::strchr("#%!Cc", s[0])) continue;  // comment lines, matutils.F90 loaddata
    // a trailing comment after the numbers (the shipped mcmcsigma2.dat: `11% example data set`)
    const size_t cpos = s.find_first_of("#%!");
    if (cpos != std::string::npos) s = s.substr(0, cpos);
    for (auto& c : s)
      if (c == ',' || c == '\t') c = ' ';
    std::istringstream ls(s);
    std::string tok;
    int n = 0;
    while (ls >> tok) {
      double d;
      if (!parse_double(tok, d)) return fail(MCMCBH_EPARSE, std::string(path) + ": not a number: " + tok);
      v.push_back(d);
      n++;
    }
    if (n == 0) continue;
    if (nc < 0) nc = n;
    else if (n != nc) return fail(MCMCBH_EPARSE, std::string(path) + ": ragged rows");
    nr++;
  }
  if (nr == 0) return fail(MCMCBH_EPARSE, std::string(path) + ": no data");
  *data = (double*)std::malloc(sizeof(double) * v.size());
  std::memcpy(*data, v.data(), sizeof(double) * v.size());
  *rows = nr;
  *cols = nc;
  return MCMCBH_OK;
}

extern "C" int mcmcbh_write_dat(const char* path, const double* x, int rows, int cols, int ld) {
  if (!path || !*path) return fail(MCMCBH_EINVAL, "empty file name");  // writedata returns stat=-1, matutils.F90:852-855
  if (!x || rows < 1 || cols < 1 || ld < rows) return fail(MCMCBH_EINVAL, "Error in writedata, empty matrix");
  FILE* fp = std::fopen(path, "w");
  if (!fp) return fail(MCMCBH_ENOFILE, std::string("Error opening file ") + path);
  for (int i = 0; i < rows; i++) {
    for (int j = 0; j < cols; j++) std::fprintf(fp, "%.17g ", x[(size_t)j * ld + i]);
    std::fputc('\n', fp);
  }
  std::fclose(fp);
  return MCMCBH_OK;
}

extern "C" int mcmcbh_write_mat4(const char* path, const char* name, const double* x, int rows, int cols, int ld) {
  if (!path || !name || !x || rows < 0 || cols < 0 || ld < rows) return fail(MCMCBH_EINVAL, "bad argument");
  FILE* fp = std::fopen(path, "wb");
  if (!fp) return fail(MCMCBH_ENOFILE, std::string("Error opening file ") + path);
  const int32_t hdr[5] = {0 /* little endian double full matrix */, rows, cols, 0, (int32_t)std::strlen(name) + 1};
  std::fwrite(hdr, sizeof hdr, 1, fp);
  std::fwrite(name, 1, std::strlen(name) + 1, fp);
  for (int j = 0; j < cols; j++) std::fwrite(x + (size_t)j * ld, sizeof(double), (size_t)rows, fp);
  const bool bad = std::ferror(fp);
  std::fclose(fp);
  return bad ? fail(MCMCBH_ENOFILE, std::string("Error writing to file ") + path) : MCMCBH_OK;
}

// addtomat4_mat (matfiles.F90:187-293): append columns to the (single) matrix of an existing MAT-v4 file --
// check the header's row count, append the new columns at the end of the file, rewrite the header with the new
// column count.  This is how the reference's 'disk' save mode streams a long chain (MCMC_aux.F90:141-160: every
// time its buffer fills it appends transpose(chain), i.e. one COLUMN per chain row).
extern "C" int mcmcbh_addto_mat4(const char* path, const double* x, int rows, int cols, int ld) {
  if (!path || !x || rows < 0 || cols < 0 || ld < rows) return fail(MCMCBH_EINVAL, "bad argument");
  FILE* fp = std::fopen(path, "r+b");
  if (!fp) return fail(MCMCBH_ENOFILE, std::string("Error opening file ") + path);
  int32_t hdr[5];
  if (std::fread(hdr, sizeof hdr, 1, fp) != 1) { std::fclose(fp); return fail(MCMCBH_EPARSE, std::string("Error reading file ") + path); }
  if (hdr[0] != 0 || hdr[3] != 0) { std::fclose(fp); return fail(MCMCBH_EPARSE, "addtomat: not a little-endian full double matrix"); }
  if (hdr[1] > 0 && hdr[1] != rows) {  // matfiles.F90:233-243
    std::fclose(fp);
    return fail(MCMCBH_EINVAL, "error: xmat should have same number of rows");
  }
  std::fseek(fp, 0, SEEK_END);
  for (int j = 0; j < cols; j++) std::fwrite(x + (size_t)j * ld, sizeof(double), (size_t)rows, fp);
  hdr[1] = rows;
  hdr[2] += cols;
  std::fseek(fp, 0, SEEK_SET);
  std::fwrite(hdr, sizeof hdr, 1, fp);
  const bool bad = std::ferror(fp);
  std::fclose(fp);
  return bad ? fail(MCMCBH_ENOFILE, std::string("Error writing new data to ") + path) : MCMCBH_OK;
}

// write_mcmcinit_namelist (mcmcinit.F90:147-179): group &mcmc with every variable of the namelist (character
// values delimited by apostrophes, as the reference's delim='APOSTROPHE'), followed by the batch group &mcmcb.
// The file reads back through mcmcbh_read_namelist (and, for &mcmc, through the reference's own reader).
extern "C" int mcmcbh_write_namelist(const char* path, const mcmcb_config* c, const mcmcbh_files* f) {
  if (!path || !c || !f) return fail(MCMCBH_EINVAL, "null argument");
  FILE* fp = std::fopen(path, "w");
  if (!fp) return fail(MCMCBH_ENOFILE, std::string("ERROR: File ") + path + " not opened for writing");
  static const char* methods[] = {"dram", "ram", "scam", "er"};
  const char* method = (c->method >= 0 && c->method <= 3) ? methods[c->method] : "dram";
  std::fprintf(fp, "&mcmc\n");
  std::fprintf(fp, " nsimu = %d,\n doadapt = %d,\n doburnin = %d,\n adaptint = %d,\n adapthist = %d,\n", c->nsimu, c->doadapt,
               c->doburnin, c->adaptint, c->adapthist);
  std::fprintf(fp, " scalelimit = %.17g,\n scalefactor = %.17g,\n drscale = %.17g,\n badaptint = %d,\n adaptend = %d,\n",
               c->scalelimit, c->scalefactor, c->drscale, c->badaptint, c->adaptend);
  std::fprintf(fp, " initcmatn = %d,\n N0 = %.17g,\n S02 = %.17g,\n filepars = %d,\n burnintime = %d,\n greedy = %d,\n",
               c->initcmatn, c->N0, c->S02, f->filepars, c->burnintime, c->greedy);
  std::fprintf(fp, " printint = %d,\n updatesigma = %d,\n usrfunlen = %d,\n", f->printint, c->updatesigma, f->usrfunlen);
  std::fprintf(fp, " chainfile = '%s',\n s2file = '%s',\n ssfile = '%s',\n condmax = %.17g,\n", f->chainfile, f->s2file, f->ssfile,
               c->condmax);
  std::fprintf(fp, " cov0file = '%s',\n covffile = '%s',\n covnfile = '%s',\n meanfile = '%s',\n nmlffile = '%s',\n", f->cov0file,
               f->covffile, f->covnfile, f->meanfile, f->nmlffile);
  std::fprintf(fp, " parfile = '%s',\n parffile = '%s',\n sigma2file = '%s',\n sigma2ffile = '%s',\n", f->parfile, f->parffile,
               f->sigma2file, f->sigma2ffile);
  std::fprintf(fp, " svddim = %d,\n condmaxini = %.17g,\n sstype = %d,\n sstrans = %.17g,\n", f->svddim, f->condmaxini, f->sstype, f->sstrans);
  std::fprintf(fp, " dumpint = %d,\n priorsfile = '%s',\n verbosity = %d,\n method = '%s',\n alphatarget = %.17g,\n nuparam = %.17g\n/\n",
               f->dumpint, f->priorsfile, f->verbosity, method, c->alphatarget, c->nuparam);
  std::fprintf(fp, "&mcmcb\n nchains = %lld,\n chain_offset = %lld,\n seed = %llu,\n device = %d,\n store_chains = %d,\n", c->nchains,
               c->chain_offset, c->seed, c->device, c->store_chains);
  std::fprintf(fp, " lanes_per_chain = %d,\n dump_stride = %d,\n kernel = %d,\n pool_adapt = %d,\n diag_stride = %d,\n diag_lags = %d,\n ngpus = %d,\n",
               c->lanes_per_chain, c->dump_stride, c->kernel, c->pool_adapt, c->diag_stride, c->diag_lags, c->ngpus);
  std::fprintf(fp, " model = '%s',\n datafile = '%s'\n/\n", c->model, f->datafile);
  const bool bad = std::ferror(fp);
  std::fclose(fp);
  return bad ? fail(MCMCBH_ENOFILE, std::string("ERROR: Error writing parameters namelist to file ") + path) : MCMCBH_OK;
}

extern "C" int mcmcbh_write_matrix(const char* path, const char* name, const double* x, int rows, int cols, int ld) {
  const size_t n = path ? std::strlen(path) : 0;
  if (n >= 4 && std::strcmp(path + n - 4, ".mat") == 0) return mcmcbh_write_mat4(path, name, x, rows, cols, ld);
  return mcmcbh_write_dat(path, x, rows, cols, ld);
}

extern "C" int mcmcbh_initialize(const char* dir, const mcmcbh_files* f, int* npar, int* nycol, double** par0, double** cmat0,
                                 double** sigma2, int** nobs, int* initcmatn) {
  if (!f || !npar || !nycol || !par0 || !cmat0 || !sigma2 || !nobs) return fail(MCMCBH_EINVAL, "null argument");
  double* m = nullptr;
  int r = 0, c = 0;
  // nycol from mcmcnycol.dat, default 1 (initialize.F90:40-50)
  *nycol = 1;
  if (mcmcbh_load_dat(join(dir, "mcmcnycol.dat").c_str(), &m, &r, &c) == 0) { *nycol = (int)m[0]; std::free(m); }
  int rc = mcmcbh_load_dat(join(dir, f->parfile).c_str(), &m, &r, &c);
  if (rc) return fail(rc, std::string("ERROR: Error reading file, ") + f->parfile);
  *npar = r * c;
  if (*npar < 1) { std::free(m); return fail(MCMCBH_EPARSE, "ERROR: npar = 0!!"); }
  *par0 = m;
  rc = mcmcbh_load_dat(join(dir, f->cov0file).c_str(), &m, &r, &c);
  if (rc || r != *npar || c != *npar) {
    if (!rc) std::free(m);
    std::free(*par0); *par0 = nullptr;
    return fail(rc ? rc : MCMCBH_EPARSE, std::string("ERROR: Error reading file ") + f->cov0file);
  }
  // row-major file order -> column-major cmat0(npar,npar)
  *cmat0 = (double*)std::malloc(sizeof(double) * (size_t)r * c);
  for (int i = 0; i < r; i++)
    for (int j = 0; j < c; j++) (*cmat0)[(size_t)j * r + i] = m[(size_t)i * c + j];
  std::free(m);
  if (initcmatn && f->covnfile[0] && mcmcbh_load_dat(join(dir, f->covnfile).c_str(), &m, &r, &c) == 0) {
    *initcmatn = (int)m[0];  // initialize.F90:93-103
    std::free(m);
  }
  *sigma2 = (double*)std::malloc(sizeof(double) * (size_t)*nycol);
  *nobs = (int*)std::malloc(sizeof(int) * (size_t)*nycol);
  if (mcmcbh_load_dat(join(dir, f->sigma2file).c_str(), &m, &r, &c) != 0) {
    for (int k = 0; k < *nycol; k++) { (*sigma2)[k] = 1.0; (*nobs)[k] = 1; }  // initialize.F90:107-110
  } else {
    // mcmcsigma2.dat is 2 x nycol: first row sigma2, second row nobs (initialize.F90:105-118)
    // the reference tests `size(s2n,1) /= 2 .and. size(s2n,2) /= nycol` (initialize.F90:111) and then indexes
    // s2n(2,1:nycol) whatever the shape; here any other shape than 2 x nycol is refused
    if (r != 2 || c != *nycol) {
      std::free(m); std::free(*par0); std::free(*cmat0); std::free(*sigma2); std::free(*nobs);
      *par0 = *cmat0 = *sigma2 = nullptr; *nobs = nullptr;
      return fail(MCMCBH_EPARSE, "ERROR: error in mcmcsigma2.dat (obs: new format 4.8.2006)");
    }
    for (int k = 0; k < *nycol; k++) { (*sigma2)[k] = m[k]; (*nobs)[k] = (int)m[(size_t)c * 1 + k]; }
    std::free(m);
  }
  return MCMCBH_OK;
}

extern "C" int mcmcbh_model_blob(const char* model, const char* datapath, double** blob, size_t* n) {
  if (!model || !datapath || !blob || !n) return fail(MCMCBH_EINVAL, "null argument");
  if (std::strcmp(model, "expreg") != 0) return fail(MCMCBH_EINVAL, std::string("no data-file loader for model ") + model);
  double* m = nullptr;
  int r = 0, c = 0;
  int rc = mcmcbh_load_dat(datapath, &m, &r, &c);
  if (rc) return rc;
  if (c < 2) { std::free(m); return fail(MCMCBH_EPARSE, "data file needs two columns (x, y)"); }
  // [n, +-max|x| (negative when some x < 0), x[npad], y[npad]] (csrc/models.cuh ExpReg)
  const int npad = (r + 1) & ~1;
  double* b = (double*)std::calloc(2 + 2 * (size_t)npad, sizeof(double));
  b[0] = r;
  bool anyneg = false;
  for (int i = 0; i < r; i++) {
    b[2 + i] = m[(size_t)i * c];
    b[2 + npad + i] = m[(size_t)i * c + 1];
    b[1] = std::max(b[1], std::fabs(m[(size_t)i * c]));
    anyneg = anyneg || m[(size_t)i * c] < 0.0;
  }
  if (anyneg) b[1] = -b[1];
  std::free(m);
  *blob = b;
  *n = 2 + 2 * (size_t)npad;
  return MCMCBH_OK;
}
