// hipims_host.cpp -- see hipims_host.h.  Host logic only; all device work goes through the C ABI.
#include "hipims_host.h"

#include <algorithm>
#include <cctype>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iterator>
#include <sstream>
#include <thread>

namespace model {
bool forceAbort = false;
std::vector<std::string> errorLog;
// src/main.cpp:631-652: warnings are logged, ModelStop sets forceAbort, Fatal ends the process
void doError(const std::string& message, unsigned char level) {
    static std::mutex mu;                              // strips report from their own host threads
    std::lock_guard<std::mutex> lock(mu);
    errorLog.push_back(message);
    const char* tag = (level & errorCodes::kLevelFatal) ? "FATAL" : (level & errorCodes::kLevelModelStop) ? "STOP" : "WARNING";
    fprintf(stderr, "[hipims %s] %s\n", tag, message.c_str());
    if (level & errorCodes::kLevelModelStop) forceAbort = true;
    if (level & errorCodes::kLevelFatal) exit(1);
}
}  // namespace model

namespace {
enum dataValues { kBedElevation = 0, kDepth, kFreeSurfaceLevel, kVelocityX, kVelocityY, kDischargeX, kDischargeY, kManningCoefficient,
                  kDisabledCells, kMaxDepth, kMaxFSL, kFroudeNumber, kUnknown = 255 };

bool isValidFloat(const char* s) {
    if (!s || !*s) return false;
    char* end = nullptr;
    strtod(s, &end);
    return end && *end == 0;
}
#define HP_CHECK(call, what)                                                                                      \
    do { if ((call) < 0) model::doError(std::string(what) + ": " + hp_last_error(), model::errorCodes::kLevelModelStop); } while (0)
}  // namespace

// ---------------------------------------------------------------------------------------------
// Util
// ---------------------------------------------------------------------------------------------
namespace Util {
double round(double dValue, unsigned char ucPlaces) {
    const unsigned int mult = static_cast<unsigned int>(std::pow(10.0, ucPlaces));
    double v = dValue * mult;
    const double rem = std::fmod(v, 1);
    v = (rem >= 0.5) ? std::ceil(v) : std::floor(v);
    return v / mult;
}
std::string toLowercase(const char* s) {
    std::string r = s ? s : "";
    std::transform(r.begin(), r.end(), r.begin(), [](unsigned char c) { return std::tolower(c); });
    return r;
}
}  // namespace Util

// ---------------------------------------------------------------------------------------------
// XML
// ---------------------------------------------------------------------------------------------
const char* XMLElement::Attribute(const char* key) const {
    for (const auto& kv : attributes) if (kv.first == key) return kv.second.c_str();
    return nullptr;
}
const XMLElement* XMLElement::FirstChildElement(const char* tag) const {
    for (const auto& c : children) if (!tag || c->name == tag) return c.get();
    return nullptr;
}
const XMLElement* XMLElement::NextSiblingElement(const char* tag) const {
    if (!parent) return nullptr;
    bool seen = false;
    for (const auto& c : parent->children) {
        if (seen && (!tag || c->name == tag)) return c.get();
        if (c.get() == this) seen = true;
    }
    return nullptr;
}

namespace {
struct XmlParser {
    const std::string& s; size_t p = 0; std::string err;
    explicit XmlParser(const std::string& text) : s(text) {}
    void skipWs() { while (p < s.size() && std::isspace(static_cast<unsigned char>(s[p]))) ++p; }
    bool starts(const char* t) const { return s.compare(p, strlen(t), t) == 0; }
    void skipMisc() {       // whitespace, comments, declarations, DOCTYPE (with an internal subset)
        for (;;) {
            skipWs();
            if (starts("<!--")) { size_t e = s.find("-->", p); p = e == std::string::npos ? s.size() : e + 3; }
            else if (starts("<?")) { size_t e = s.find("?>", p); p = e == std::string::npos ? s.size() : e + 2; }
            else if (starts("<!")) {
                int depth = 0;
                while (p < s.size()) { char c = s[p++]; if (c == '[') ++depth; else if (c == ']') --depth; else if (c == '>' && depth <= 0) break; }
            } else return;
        }
    }
    static std::string unescape(const std::string& v) {
        std::string r; r.reserve(v.size());
        for (size_t i = 0; i < v.size(); ++i) {
            if (v[i] != '&') { r += v[i]; continue; }
            if (v.compare(i, 5, "&amp;") == 0) { r += '&'; i += 4; } else if (v.compare(i, 4, "&lt;") == 0) { r += '<'; i += 3; }
            else if (v.compare(i, 4, "&gt;") == 0) { r += '>'; i += 3; } else if (v.compare(i, 6, "&quot;") == 0) { r += '"'; i += 5; }
            else if (v.compare(i, 6, "&apos;") == 0) { r += '\''; i += 5; } else r += v[i];
        }
        return r;
    }
    std::string name() { size_t b = p; while (p < s.size() && (std::isalnum(static_cast<unsigned char>(s[p])) || strchr("_-:.", s[p]))) ++p; return s.substr(b, p - b); }
    std::unique_ptr<XMLElement> element(const XMLElement* parent) {
        if (p >= s.size() || s[p] != '<') { err = "expected '<'"; return nullptr; }
        ++p;
        auto e = std::make_unique<XMLElement>();
        e->parent = parent; e->name = name();
        if (e->name.empty()) { err = "empty element name"; return nullptr; }
        for (;;) {
            skipWs();
            if (p >= s.size()) { err = "unexpected end inside <" + e->name + ">"; return nullptr; }
            if (starts("/>")) { p += 2; return e; }
            if (s[p] == '>') { ++p; break; }
            std::string key = name();
            skipWs();
            if (key.empty() || p >= s.size() || s[p] != '=') { err = "bad attribute in <" + e->name + ">"; return nullptr; }
            ++p; skipWs();
            const char q = s[p];
            if (q != '"' && q != '\'') { err = "unquoted attribute in <" + e->name + ">"; return nullptr; }
            size_t end = s.find(q, p + 1);
            if (end == std::string::npos) { err = "unterminated attribute"; return nullptr; }
            e->attributes.emplace_back(key, unescape(s.substr(p + 1, end - p - 1)));
            p = end + 1;
        }
        for (;;) {
            size_t lt = s.find('<', p);
            if (lt == std::string::npos) { err = "missing </" + e->name + ">"; return nullptr; }
            e->text += unescape(s.substr(p, lt - p));
            p = lt;
            if (starts("</")) {
                p += 2; std::string n = name(); skipWs();
                if (n != e->name || p >= s.size() || s[p] != '>') { err = "mismatched </" + n + ">"; return nullptr; }
                ++p; return e;
            }
            if (starts("<!--") || starts("<?") || starts("<!")) { skipMisc(); continue; }
            auto child = element(e.get());
            if (!child) return nullptr;
            e->children.push_back(std::move(child));
        }
    }
};
}  // namespace

bool XMLDocument::Parse(const std::string& xml) {
    XmlParser ps(xml);
    ps.skipMisc();
    root = ps.element(nullptr);
    error = ps.err;
    return root != nullptr;
}
bool XMLDocument::LoadFile(const std::string& path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) { error = "cannot open " + path; return false; }
    std::stringstream ss; ss << f.rdbuf();
    return Parse(ss.str());
}

// ---------------------------------------------------------------------------------------------
// CSV, raster
// ---------------------------------------------------------------------------------------------
bool CCSVDataset::readFile() {
    std::ifstream f(sFilename);
    if (!f) return false;
    std::string line;
    while (std::getline(f, line)) {
        while (!line.empty() && (line.back() == '\r' || line.back() == '\n' || line.back() == ' ')) line.pop_back();
        if (line.empty()) continue;
        std::vector<std::string> cells; std::string cell; std::stringstream ss(line);
        while (std::getline(ss, cell, ',')) {
            size_t b = cell.find_first_not_of(" \t"), e = cell.find_last_not_of(" \t");
            cells.push_back(b == std::string::npos ? "" : cell.substr(b, e - b + 1));
        }
        rows.push_back(cells);
    }
    bReady = true;
    return true;
}

bool SRaster::read(const std::string& path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) return false;
    char tag[16] = {0};
    f.read(tag, 15);
    f.close();
    return std::string(tag) == "EHFA_HEADER_TAG" ? readHFA(path) : readASCII(path);
}

// ERDAS IMAGINE HFA: header tag -> Ehfa_File -> tree of Ehfa_Entry nodes; the first Eimg_Layer gives
// width/height/pixel type/block size, its RasterDMS (Edms_State) child the block table, its Map_Info child
// the georeferencing.  Compressed blocks: {u32 min, i32 runs, i32 data offset, u8 bits} then run lengths
// (top two bits of the first byte = number of extra bytes, big-endian) and values (big-endian, offset by min).
bool SRaster::readHFA(const std::string& path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) return false;
    std::vector<unsigned char> d((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    auto u32 = [&](size_t o) -> uint32_t { return o + 4 <= d.size() ? (uint32_t)d[o] | (uint32_t)d[o + 1] << 8 | (uint32_t)d[o + 2] << 16 | (uint32_t)d[o + 3] << 24 : 0u; };
    auto u16 = [&](size_t o) -> uint32_t { return o + 2 <= d.size() ? (uint32_t)d[o] | (uint32_t)d[o + 1] << 8 : 0u; };
    auto f64 = [&](size_t o) { double v = 0; if (o + 8 <= d.size()) memcpy(&v, &d[o], 8); return v; };
    if (d.size() < 64) return false;
    const size_t hdr = u32(16), root = u32(hdr + 8);
    struct Entry { size_t next, child, data, size; std::string name, type; };
    auto entry = [&](size_t o) {
        Entry e{u32(o), u32(o + 12), u32(o + 16), u32(o + 20), "", ""};
        if (o + 120 <= d.size()) { e.name = reinterpret_cast<const char*>(&d[o + 24]); e.type = std::string(reinterpret_cast<const char*>(&d[o + 88]), strnlen(reinterpret_cast<const char*>(&d[o + 88]), 32)); }
        return e;
    };
    size_t layer = 0;
    for (size_t o = entry(root).child; o && o < d.size(); o = entry(o).next) if (entry(o).type == "Eimg_Layer") { layer = o; break; }
    if (!layer) return false;
    const Entry L = entry(layer);
    const long W = (int32_t)u32(L.data), H = (int32_t)u32(L.data + 4), BWd = (int32_t)u32(L.data + 12), BHt = (int32_t)u32(L.data + 16);
    const unsigned ptype = u16(L.data + 10);     // u1,u2,u4,u8,s8,u16,s16,u32,s32,f32,f64
    if (W <= 0 || H <= 0 || BWd <= 0 || BHt <= 0 || ptype < 3 || ptype > 10) return false;
    const int psize = (ptype == 3 || ptype == 4) ? 1 : (ptype == 5 || ptype == 6) ? 2 : (ptype == 10 ? 8 : 4);
    size_t dms = 0, mapinfo = 0;
    for (size_t o = L.child; o && o < d.size(); o = entry(o).next) { const Entry e = entry(o); if (e.type == "Edms_State") dms = e.data; if (e.type == "Eprj_MapInfo") mapinfo = e.data; }
    if (!dms) return false;
    const uint32_t nblocks = u32(dms + 14);
    const size_t table = u32(dms + 18);
    const long nbx = (W + BWd - 1) / BWd, nby = (H + BHt - 1) / BHt;
    if ((long)nblocks < nbx * nby) return false;
    cols = W; rows = H; nodata = -9999.0;
    values.assign((size_t)W * H, nodata);
    auto convert = [&](uint32_t bits_lo, const unsigned char* raw) -> double {   // raw: little-endian pixel (uncompressed blocks)
        (void)bits_lo;
        switch (ptype) {
        case 3: return raw[0]; case 4: return (int8_t)raw[0];
        case 5: return (uint16_t)(raw[0] | raw[1] << 8); case 6: return (int16_t)(raw[0] | raw[1] << 8);
        case 7: { uint32_t v; memcpy(&v, raw, 4); return v; } case 8: { int32_t v; memcpy(&v, raw, 4); return v; }
        case 9: { float v; memcpy(&v, raw, 4); return v; } default: { double v; memcpy(&v, raw, 8); return v; }
        }
    };
    std::vector<double> blk((size_t)BWd * BHt);
    for (long b = 0; b < nbx * nby; ++b) {
        const size_t rec = table + (size_t)b * 14;
        const size_t off = u32(rec + 2), size = u32(rec + 6);
        const unsigned comp = u16(rec + 12);
        if (off + size > d.size()) return false;
        const unsigned char* c = &d[off];
        if (comp == 0) {
            if (size < blk.size() * psize) return false;
            for (size_t i = 0; i < blk.size(); ++i) blk[i] = convert(0, c + i * psize);
        } else {
            if (psize == 8) return false;                 // compressed f64 is not produced by the tools in use
            const uint32_t vmin = u32(off); const int32_t runs = (int32_t)u32(off + 4); const uint32_t doff = u32(off + 8); const unsigned nbits = d[off + 12];
            auto value = [&](const unsigned char* v, size_t i) -> uint32_t {
                switch (nbits) {
                case 0: return 0; case 8: return v[i]; case 16: return (uint32_t)v[2 * i] << 8 | v[2 * i + 1];
                case 32: return (uint32_t)v[4 * i] << 24 | (uint32_t)v[4 * i + 1] << 16 | (uint32_t)v[4 * i + 2] << 8 | v[4 * i + 3];
                case 1: return (v[i >> 3] >> (i & 7)) & 1; case 2: return (v[i >> 2] >> ((i & 3) * 2)) & 3; case 4: return (v[i >> 1] >> ((i & 1) * 4)) & 15;
                default: return 0;
                }
            };
            auto store = [&](size_t i, uint32_t raw) {
                raw += vmin;
                if (ptype == 9) { float fv; memcpy(&fv, &raw, 4); blk[i] = fv; }
                else if (ptype == 4) blk[i] = (int8_t)raw; else if (ptype == 6) blk[i] = (int16_t)raw; else if (ptype == 8) blk[i] = (int32_t)raw; else blk[i] = raw;
            };
            if (runs == -1) {
                for (size_t i = 0; i < blk.size(); ++i) store(i, value(c + 13, i));
            } else {
                size_t cnt = 13, n = 0;
                for (int32_t r = 0; r < runs; ++r) {
                    if (cnt >= size) return false;
                    const unsigned extra = c[cnt] >> 6;
                    size_t len = c[cnt] & 0x3f;
                    for (unsigned j = 0; j < extra; ++j) len = len << 8 | c[cnt + 1 + j];
                    cnt += 1 + extra;
                    const uint32_t v = value(c + doff, (size_t)r);
                    for (size_t j = 0; j < len && n < blk.size(); ++j) store(n++, v);
                }
                if (n != blk.size()) return false;
            }
        }
        const long by = b / nbx, bx = b % nbx;
        for (long y = 0; y < BHt; ++y) for (long x = 0; x < BWd; ++x) {
            const long gy = by * BHt + y, gx = bx * BWd + x;                    // file rows are north-first
            if (gy < H && gx < W) values[(size_t)(H - 1 - gy) * W + gx] = blk[(size_t)y * BWd + x];
        }
    }
    cellsize = 1.0; xll = yll = 0.0;
    if (mapinfo) {          // Eprj_MapInfo: name, upperLeftCenter, lowerRightCenter, pixelSize, units -- each behind {count, ptr}
        size_t o = mapinfo; o += 8 + u32(o);
        const double ulx = f64(o + 8); o += 24;
        const double lry = f64(o + 16); o += 24;
        cellsize = f64(o + 8);
        xll = ulx - 0.5 * cellsize; yll = lry - 0.5 * cellsize;
    }
    return true;
}

bool SRaster::readASCII(const std::string& path) {
    std::ifstream f(path);
    if (!f) return false;
    std::string key; double v;
    std::map<std::string, double> hdr;
    for (int i = 0; i < 6; ++i) {
        std::streampos pos = f.tellg();
        if (!(f >> key)) return false;
        if (!std::isalpha(static_cast<unsigned char>(key[0]))) { f.seekg(pos); break; }
        if (!(f >> v)) return false;
        hdr[Util::toLowercase(key.c_str())] = v;
    }
    cols = static_cast<unsigned long>(hdr["ncols"]); rows = static_cast<unsigned long>(hdr["nrows"]);
    cellsize = hdr.count("cellsize") ? hdr["cellsize"] : 1.0;
    xll = hdr.count("xllcorner") ? hdr["xllcorner"] : hdr["xllcenter"]; yll = hdr.count("yllcorner") ? hdr["yllcorner"] : hdr["yllcenter"];
    nodata = hdr.count("nodata_value") ? hdr["nodata_value"] : -9999.0;
    if (!cols || !rows) return false;
    values.assign(cols * rows, nodata);
    for (unsigned long r = 0; r < rows; ++r)            // file is north-first, arrays are south-first
        for (unsigned long c = 0; c < cols; ++c) if (!(f >> values[(rows - 1 - r) * cols + c])) return false;
    return true;
}
bool SRaster::write(const std::string& path) const {
    FILE* f = fopen(path.c_str(), "w");
    if (!f) return false;
    fprintf(f, "ncols %lu\nnrows %lu\nxllcorner %.10g\nyllcorner %.10g\ncellsize %.10g\nNODATA_value %.10g\n", cols, rows, xll, yll, cellsize, nodata);
    for (unsigned long r = 0; r < rows; ++r) {
        for (unsigned long c = 0; c < cols; ++c) fprintf(f, c ? " %.17g" : "%.17g", values[(rows - 1 - r) * cols + c]);
        fputc('\n', f);
    }
    fclose(f);
    return true;
}

// ---------------------------------------------------------------------------------------------
// Executor
// ---------------------------------------------------------------------------------------------
CExecutorControlCUDA* CExecutorControlCUDA::createFromConfig(const XMLElement* pXExecution) {
    const XMLElement* pX = pXExecution ? pXExecution->FirstChildElement("executor") : nullptr;
    const std::string name = Util::toLowercase(pX ? pX->Attribute("name") : "cuda");
    // "OpenCL" configurations are accepted: the CUDA executor is the drop-in for that slot
    if (name != "cuda" && name != "opencl") {
        model::doError("Unsupported executor specified in configuration.", model::errorCodes::kLevelFatal);
        return nullptr;
    }
    auto* ex = new CExecutorControlCUDA();
    if (!ex->setupFromConfig(pX)) { delete ex; return nullptr; }
    return ex;
}
bool CExecutorControlCUDA::setupFromConfig(const XMLElement* pX) {
    int n = 0;
    if (hp_device_count(&n) < 0) { model::doError(std::string("No CUDA device: ") + hp_last_error(), model::errorCodes::kLevelModelStop); return false; }
    uiDeviceCount = static_cast<unsigned int>(n);
    int device = 1;   // 1-based like getDevice(n)
    for (const XMLElement* p = pX ? pX->FirstChildElement("parameter") : nullptr; p; p = p->NextSiblingElement("parameter")) {
        const std::string key = Util::toLowercase(p->Attribute("name"));
        if (key == "devicenumber" && p->Attribute("value")) device = atoi(p->Attribute("value"));
        // deviceFilter (GPU|CPU|APU) is accepted and ignored: there is exactly one kind of device here
    }
    if (device < 1 || device > n) { model::doError("Invalid device number in configuration.", model::errorCodes::kLevelModelStop); return false; }
    if (hp_executor_create(device - 1, nullptr, &pExecutor) < 0) { model::doError(hp_last_error(), model::errorCodes::kLevelModelStop); return false; }
    iDeviceOrdinal = device - 1;
    char buf[256]; int sms = 0; size_t mem = 0;
    hp_executor_describe(pExecutor, buf, sizeof(buf), &sms, &mem);
    sDeviceName = buf;
    return true;
}
CExecutorControlCUDA::~CExecutorControlCUDA() { if (pExecutor) hp_executor_destroy(pExecutor); }
void CExecutorControlCUDA::blockUntilFinished() { if (pExecutor) hp_executor_finish(pExecutor); }

// ---------------------------------------------------------------------------------------------
// Boundaries
// ---------------------------------------------------------------------------------------------
namespace {
// shared by the uniform and cell series: header row skipped, N numeric columns
bool readSeries(CCSVDataset* csv, size_t columns, std::vector<double>& out, double& interval, double& length) {
    bool invalid = false, header = false;
    out.clear();
    for (const auto& row : csv->rows) {
        if (!header) { header = true; continue; }                 // first row is always treated as headers
        if (row.size() == columns) {
            for (const auto& c : row) { if (!isValidFloat(c.c_str())) invalid = true; out.push_back(atof(c.c_str())); }
        } else { invalid = true; for (size_t i = 0; i < columns; ++i) out.push_back(0.0); }
    }
    if (invalid) model::doError("Some CSV entries were not valid for a boundary timeseries.", model::errorCodes::kLevelWarning);
    const size_t n = out.size() / columns;
    if (n < 2) { model::doError("A boundary timeseries is too short.", model::errorCodes::kLevelWarning); return false; }
    interval = out[columns] - out[0];
    length = out[(n - 1) * columns];
    return true;
}
}  // namespace

bool CBoundaryUniform::setupFromConfig(const XMLElement* pElement, const std::string& dir) {
    sName = pElement->Attribute("name") ? pElement->Attribute("name") : "";
    const std::string value = Util::toLowercase(pElement->Attribute("value"));
    if (value.empty() || value == "rain-intensity") ucValue = 0;
    else if (value == "loss-rate") ucValue = 1;
    else model::doError("Unrecognised value for uniform timeseries file.", model::errorCodes::kLevelWarning);
    CCSVDataset csv(dir + Util::toLowercase(pElement->Attribute("source")));
    if (!csv.readFile()) { model::doError("Could not read a uniform boundary timeseries file.", model::errorCodes::kLevelWarning); return false; }
    importTimeseries(&csv);
    return true;
}
void CBoundaryUniform::importTimeseries(CCSVDataset* csv) { readSeries(csv, 2, series, dTimeseriesInterval, dTimeseriesLength); }
void CBoundaryUniform::prepareBoundary(hp_scheme* s, CDomainCartesian*, double) {
    if (series.size() < 4) return;
    hp_bdy_uniform conf{static_cast<uint32_t>(series.size() / 2), ucValue, dTimeseriesInterval, dTimeseriesLength};
    HP_CHECK(hp_boundary_add_uniform(s, &conf, series.data()), "uniform boundary " + sName);
}

bool CBoundaryCell::setupFromConfig(const XMLElement* pElement, const std::string& dir) {
    sName = pElement->Attribute("name") ? pElement->Attribute("name") : "";
    const char* dis = pElement->Attribute("dischargeValue");
    const char* dep = pElement->Attribute("depthValue");
    const std::string d = Util::toLowercase(dis), h = Util::toLowercase(dep);
    bDischargeIsTotal = false;
    if (!dis || d == "total") { ucDischargeValue = 1; bDischargeIsTotal = true; }
    else if (d == "cell") { ucDischargeValue = 1; bDischargeIsTotal = true; }     // reference: same enum value as "total"
    else if (d == "velocity") ucDischargeValue = 2;
    else if (d == "ignore" || d == "disabled") ucDischargeValue = 0;
    else if (d == "volume" || d == "surging") ucDischargeValue = 3;
    else model::doError("Unrecognised discharge parameter specified for timeseries file.", model::errorCodes::kLevelWarning);
    if (!dep || h == "fsl") ucDepthValue = 1;
    else if (h == "depth") ucDepthValue = 2;
    else if (h == "ignore" || h == "disabled") ucDepthValue = 0;
    else model::doError("Unrecognised depth parameter specified in timeseries file.", model::errorCodes::kLevelWarning);
    CCSVDataset csv(dir + Util::toLowercase(pElement->Attribute("source")));
    if (!csv.readFile()) { model::doError("Could not read a boundary timeseries file.", model::errorCodes::kLevelWarning); return false; }
    importTimeseries(&csv);
    if (pElement->Attribute("mapFile")) {
        CCSVDataset map(dir + Util::toLowercase(pElement->Attribute("mapFile")));
        if (!map.readFile()) { model::doError("Could not read a boundary map file.", model::errorCodes::kLevelWarning); return false; }
        importMap(&map);
    }
    return true;
}
void CBoundaryCell::importTimeseries(CCSVDataset* csv) { readSeries(csv, 4, series, dTimeseriesInterval, dTimeseriesLength); }
void CBoundaryCell::importMap(CCSVDataset* csv) {
    bool header = false, invalid = false;
    for (const auto& row : csv->rows) {
        if (!header) { header = true; continue; }
        if (row.size() == 2 || (row.size() == 3 && row[2] == sName)) relations.emplace_back(atoi(row[0].c_str()), atoi(row[1].c_str()));
        else if (row.size() != 3) invalid = true;
    }
    if (invalid) model::doError("Some CSV entries were not valid for a boundary map file.", model::errorCodes::kLevelWarning);
}
void CBoundaryCell::prepareBoundary(hp_scheme* s, CDomainCartesian* pDomain, double) {
    if (series.size() < 8 || relations.empty()) return;
    std::vector<double> ts = series;
    if (bDischargeIsTotal)                               // CBoundaryCell.cpp:352-356: Q divided by the number of mapped cells
        for (size_t i = 0; i < ts.size() / 4; ++i) { ts[4 * i + 2] /= relations.size(); ts[4 * i + 3] /= relations.size(); }
    std::vector<uint64_t> ids;
    for (const auto& r : relations) ids.push_back(pDomain->getCellID(r.first, r.second));
    hp_bdy_cell conf{ts.size() / 4, dTimeseriesInterval, dTimeseriesLength, ids.size(), ucDepthValue, ucDischargeValue};
    HP_CHECK(hp_boundary_add_cell(s, &conf, ids.data(), ts.data()), "cell boundary " + sName);
}

bool CBoundaryGridded::setupFromConfig(const XMLElement* pElement, const std::string& dir) {
    sName = pElement->Attribute("name") ? pElement->Attribute("name") : "";
    sMask = pElement->Attribute("mask") ? pElement->Attribute("mask") : "";
    sSourceDir = dir;
    if (!isValidFloat(pElement->Attribute("interval"))) { model::doError("Gridded boundary interval is not a valid number.", model::errorCodes::kLevelWarning); return false; }
    dInterval = atof(pElement->Attribute("interval"));
    const std::string value = Util::toLowercase(pElement->Attribute("value"));
    if (value.empty() || value == "rain-intensity") ucValue = 0;
    else if (value == "mass-flux") ucValue = 2;           // the reference's host maps this to 1, which its kernel ignores
    else model::doError("Unrecognised value parameter specified for gridded timeseries data.", model::errorCodes::kLevelWarning);
    return true;
}
void CBoundaryGridded::prepareBoundary(hp_scheme* s, CDomainCartesian* pDomain, double dSimulationLength) {
    // frames: mask with %n replaced by the frame index (the reference formats a timestamp through
    // GDAL-readable rasters; here ESRI ASCII grids named by index)
    std::vector<double> frames; SRaster first; uint64_t entries = 0;
    for (double t = 0.0; t <= dSimulationLength; t += dInterval) {
        std::string file = sMask; const size_t pos = file.find("%n");
        if (pos != std::string::npos) file.replace(pos, 2, std::to_string(entries));
        SRaster r;
        if (!r.read(sSourceDir + file)) { model::doError("Gridded boundary raster missing for frame " + std::to_string(entries), model::errorCodes::kLevelWarning); break; }
        if (entries == 0) first = r;
        if (r.cols != first.cols || r.rows != first.rows) { model::doError("Gridded boundary rasters differ in size.", model::errorCodes::kLevelWarning); break; }
        frames.insert(frames.end(), r.values.begin(), r.values.end());
        ++entries;
    }
    if (!entries) return;
    // the kernel subtracts the offset of the coarse grid's south-west corner from the cell position
    hp_bdy_gridded conf{dInterval, first.cellsize, first.xll - pDomain->dRealOffsetX, first.yll - pDomain->dRealOffsetY, entries, ucValue, first.rows, first.cols};
    HP_CHECK(hp_boundary_add_gridded(s, &conf, frames.data()), "gridded boundary " + sName);
}

bool CBoundaryMap::setupFromConfig(const XMLElement* pX, const std::string& sConfigDir) {
    if (!pX) return true;
    std::string dir = pX->Attribute("sourceDir") ? pX->Attribute("sourceDir") : "";
    if (dir.empty() || dir[0] != '/') dir = sConfigDir + dir;
    // <domainEdge> elements are accepted and ignored, as in the reference
    unsigned int autoName = 0;
    for (const XMLElement* t = pX->FirstChildElement("timeseries"); t; t = t->NextSiblingElement("timeseries")) {
        const std::string type = Util::toLowercase(t->Attribute("type"));
        std::unique_ptr<CBoundary> b;
        if (type == "cell") b.reset(new CBoundaryCell());
        else if (type == "atmospheric" || type == "uniform") b.reset(new CBoundaryUniform());
        else if (type == "gridded" || type == "spatially-varying") b.reset(new CBoundaryGridded());
        else { model::doError("Ignored boundary timeseries of unrecognised type.", model::errorCodes::kLevelWarning); continue; }
        XMLElement withName;
        const XMLElement* use = t;
        if (!t->Attribute("name")) { withName.attributes = t->attributes; withName.attributes.emplace_back("name", "Boundary_" + std::to_string(++autoName)); use = &withName; }
        if (!b->setupFromConfig(use, dir)) { model::doError("Encountered an error loading a boundary definition.", model::errorCodes::kLevelWarning); continue; }
        boundaries.push_back(std::move(b));
    }
    return true;
}
void CBoundaryMap::prepareBoundaries(hp_scheme* s, CDomainCartesian* d, double len) { for (auto& b : boundaries) b->prepareBoundary(s, d, len); }
CBoundary* CBoundaryMap::getBoundaryByName(const std::string& n) { for (auto& b : boundaries) if (b->getName() == n) return b.get(); return nullptr; }

// ---------------------------------------------------------------------------------------------
// Domain
// ---------------------------------------------------------------------------------------------
unsigned char CDomainCartesian::getDataValueCode(const std::string& v) {   // src/Domain/CDomain.cpp:464-500
    auto has = [&](const char* k) { return v.find(k) != std::string::npos; };
    if (has("dem")) return kBedElevation;
    if (has("maxdepth")) return kMaxDepth; else if (has("depth")) return kDepth;
    if (has("disabled")) return kDisabledCells;
    if (has("dischargex")) return kDischargeX;
    if (has("dischargey")) return kDischargeY;
    if (has("maxfsl")) return kMaxFSL; else if (has("fsl")) return kFreeSurfaceLevel;
    if (has("manningcoefficient")) return kManningCoefficient;
    if (has("velocityx")) return kVelocityX;
    if (has("velocityy")) return kVelocityY;
    if (has("froude")) return kFroudeNumber;
    return kUnknown;
}

void CDomainCartesian::handleInputData(unsigned long id, double v, unsigned char code, unsigned char rounding) {   // CDomain.cpp:294-397
    double* st = &dCellStates[4 * id];
    switch (code) {
    case kBedElevation: dBedElevations[id] = Util::round(v, rounding); st[0] = Util::round(v, rounding); break;
    case kFreeSurfaceLevel: st[0] = Util::round(v, rounding); st[1] = Util::round(v, rounding); break;
    case kDepth: st[0] = Util::round(dBedElevations[id] + v, rounding); st[1] = st[0]; break;
    case kDisabledCells: if (v > 1.0 && v < 9999.0) st[1] = Util::round(-9999.0, rounding); break;
    case kDischargeX: st[2] = Util::round(v, rounding); break;
    case kDischargeY: st[3] = Util::round(v, rounding); break;
    case kVelocityX: st[2] = Util::round(v * (st[0] - dBedElevations[id]), rounding); break;
    case kVelocityY: st[3] = Util::round(v * (st[0] - dBedElevations[id]), rounding); break;
    case kManningCoefficient: dManningValues[id] = Util::round(v, rounding); break;
    default: break;
    }
}

bool CDomainCartesian::loadInitialConditionSource(unsigned char code, const std::string& type, const std::string& source) {
    if (type == "constant") {
        if (!isValidFloat(source.c_str())) { model::doError("Invalid source constant given.", model::errorCodes::kLevelWarning); return false; }
        const double v = atof(source.c_str());
        for (unsigned long i = 0; i < getCellCount(); ++i) handleInputData(i, v, code, 4);
        return true;
    }
    if (type == "raster") {
        SRaster r;
        if (!r.read(sSourceDir + source) || r.cols != ulCols || r.rows != ulRows) { model::doError("Raster source could not be read or does not match the domain.", model::errorCodes::kLevelWarning); return false; }
        for (unsigned long i = 0; i < getCellCount(); ++i) handleInputData(i, r.values[i], code, 4);
        return true;
    }
    model::doError("Unrecognised data source type.", model::errorCodes::kLevelWarning);
    return false;
}

bool CDomainCartesian::configureDomain(const XMLElement* pXDomain, const std::string& sConfigDir) {
    const XMLElement* pXData = pXDomain->FirstChildElement("data");
    if (!pXData) { model::doError("The <domain> element has no <data> element.", model::errorCodes::kLevelModelStop); return false; }
    auto dirOf = [&](const char* a) { std::string d = pXData->Attribute(a) ? pXData->Attribute(a) : ""; if (!d.empty() && d[0] != '/') d = sConfigDir + d; return d; };
    sSourceDir = dirOf("sourceDir"); sTargetDir = dirOf("targetDir");
    // 1. structure: the raster tagged "structure" fixes cols/rows/resolution (CDomainCartesian.cpp:69-160)
    struct Src { std::string type, value, source; unsigned char code; };
    std::vector<Src> others; Src dem{"", "", "", kUnknown}, depth{"", "", "", kUnknown};
    for (const XMLElement* s = pXData->FirstChildElement("dataSource"); s; s = s->NextSiblingElement("dataSource")) {
        Src src{Util::toLowercase(s->Attribute("type")), Util::toLowercase(s->Attribute("value")), s->Attribute("source") ? s->Attribute("source") : "", kUnknown};
        src.code = getDataValueCode(src.value);
        if (src.value.find("structure") != std::string::npos) {
            SRaster r;
            if (!r.read(sSourceDir + src.source)) { model::doError("Could not open the domain structure raster.", model::errorCodes::kLevelModelStop); return false; }
            ulCols = r.cols; ulRows = r.rows; dCellResolution = r.cellsize; dRealOffsetX = r.xll; dRealOffsetY = r.yll;
        }
        if (src.code == kBedElevation) dem = src;
        else if (src.code == kDepth || src.code == kFreeSurfaceLevel) depth = src;
        else others.push_back(src);
    }
    if (!ulCols || !ulRows) { model::doError("No structure source defined for the domain.", model::errorCodes::kLevelModelStop); return false; }
    dCellStates.assign(4 * getCellCount(), 0.0); dBedElevations.assign(getCellCount(), 0.0); dManningValues.assign(getCellCount(), 0.0);
    if (dem.code == kUnknown || depth.code == kUnknown) model::doError("Missing DEM or depth data source.", model::errorCodes::kLevelWarning);
    // 2. initial conditions in the order DEM, depth/FSL, everything else (CDomainCartesian.cpp:252-283)
    if (dem.code != kUnknown && !loadInitialConditionSource(dem.code, dem.type, dem.source)) { model::doError("Could not load DEM data.", model::errorCodes::kLevelWarning); return false; }
    if (depth.code != kUnknown && !loadInitialConditionSource(depth.code, depth.type, depth.source)) { model::doError("Could not load depth/FSL data.", model::errorCodes::kLevelWarning); return false; }
    for (const auto& o : others) if (o.code != kUnknown && !loadInitialConditionSource(o.code, o.type, o.source)) { model::doError("Could not load initial conditions.", model::errorCodes::kLevelWarning); return false; }
    // 3. outputs (CDomainCartesian.cpp:288-339)
    for (const XMLElement* t = pXData->FirstChildElement("dataTarget"); t; t = t->NextSiblingElement("dataTarget"))
        outputs.push_back({Util::toLowercase(t->Attribute("value")), t->Attribute("format") ? t->Attribute("format") : "", t->Attribute("target") ? t->Attribute("target") : ""});
    // 4. boundaries (sourceDir is relative to the configuration file like the data dirs)
    return boundaryMap.setupFromConfig(pXDomain->FirstChildElement("boundaryConditions"), sConfigDir);
}

double CDomainCartesian::getVolume() const {
    double v = 0.0;
    for (unsigned long i = 0; i < getCellCount(); ++i) {
        if (dCellStates[4 * i + 1] <= -9999.0 || dBedElevations[i] <= -9999.0) continue;
        v += std::max(0.0, dCellStates[4 * i] - dBedElevations[i]) * dCellResolution * dCellResolution;
    }
    return v;
}

// src/Datasets/CRasterDataset.cpp:185-267
double CDomainCartesian::deriveOutput(unsigned char code, const double* st, double bed, double res, double nodata) {
    const double depth = st[0] - bed;
    switch (code) {
    case kMaxFSL: return (st[1] < bed + 1E-8 || bed > 9999.0) ? nodata : st[1];
    case kFreeSurfaceLevel: return (st[0] < bed + 1E-8 || bed > 9999.0) ? nodata : st[0];
    case kMaxDepth: { const double d = std::max(0.0, st[1] - bed); return (d < 1E-8 || d <= -9990.0 || d >= 9999.0) ? nodata : d; }
    case kDepth: { const double d = std::max(0.0, depth); return d < 1E-8 ? nodata : d; }
    case kDischargeX: return st[2] * res;
    case kDischargeY: return st[3] * res;
    case kVelocityX: return depth > 1E-8 ? st[2] / depth : nodata;
    case kVelocityY: return depth > 1E-8 ? st[3] / depth : nodata;
    case kFroudeNumber: { const double u = st[2] / depth, v = st[3] / depth; return depth > 1E-8 ? std::sqrt(u * u + v * v) / std::sqrt(9.81 * depth) : nodata; }
    default: return nodata;
    }
}

bool CDomainCartesian::writeOutputs(double dTime, CScheme* pScheme) {       // src/Domain/Cartesian/CDomainCartesian.cpp:738-767
    bool ok = true;
    std::vector<double> values;
    for (const auto& o : outputs) {
        const unsigned char code = getDataValueCode(o.sValue);
        if (!(pScheme && pScheme->deriveRaster(code, values))) {
            values.resize(getCellCount());
            for (unsigned long y = 0; y < ulRows; ++y)
                for (unsigned long x = 0; x < ulCols; ++x) {
                    const unsigned long i = getCellID(x, y);
                    values[(ulRows - 1 - y) * ulCols + x] = deriveOutput(code, &dCellStates[4 * i], dBedElevations[i], dCellResolution, -9999.0);
                }
        }
        std::string file = o.sTarget; const size_t pos = file.find("%t");
        char tbuf[64]; snprintf(tbuf, sizeof(tbuf), "%g", std::floor(dTime * 100.0) / 100.0);   // "%t" -> floor(t*100)/100
        if (pos != std::string::npos) file.replace(pos, 2, tbuf);
        ok = CRasterDataset::writeRaster(o.sFormat, sTargetDir + file, ulCols, ulRows, dRealOffsetX, dRealOffsetY, dCellResolution, values.data()) && ok;
    }
    return ok;
}

// ---------------------------------------------------------------------------------------------
// Multi-domain sets -> one domain
// ---------------------------------------------------------------------------------------------
CDomainCartesian* CDomainCartesian::mergeStacked(std::vector<std::unique_ptr<CDomainCartesian>>& parts, std::vector<unsigned long>& off) {
    const size_t n = parts.size();
    std::vector<size_t> order(n);
    for (size_t i = 0; i < n; ++i) order[i] = i;
    std::sort(order.begin(), order.end(), [&](size_t a, size_t b) { return parts[a]->dRealOffsetY < parts[b]->dRealOffsetY; });
    const CDomainCartesian& first = *parts[order[0]];
    const double res = first.dCellResolution;
    off.assign(n, 0);
    std::vector<unsigned long> split(n + 1, 0);          // merged row where part order[k] starts to be authoritative
    unsigned long total = first.ulRows;
    for (size_t k = 1; k < n; ++k) {
        const CDomainCartesian &lo = *parts[order[k - 1]], &hi = *parts[order[k]];
        // the linking rules of CDomainLink::canLink (Links/CDomainLink.cpp:73-136)
        if (hi.dCellResolution != res) { model::doError("Cannot merge domains of mismatched resolutions.", model::errorCodes::kLevelModelStop); return nullptr; }
        if (hi.ulCols != first.ulCols || std::fabs(hi.dRealOffsetX - first.dRealOffsetX) > 0.1 * res) {
            model::doError("Only domains stacked north-south over the same columns can be merged.", model::errorCodes::kLevelModelStop); return nullptr; }
        const double top = lo.dRealOffsetY + res * lo.ulRows;
        const double gap = top - hi.dRealOffsetY;       // overlap in metres
        if (gap < -0.1 * res) { model::doError("Domains do not overlap or touch on the N/S axis.", model::errorCodes::kLevelModelStop); return nullptr; }
        if (std::fabs(std::remainder(gap, res)) > 0.1 * res) { model::doError("Cannot merge domains that are not aligned N/S.", model::errorCodes::kLevelModelStop); return nullptr; }
        const unsigned long ov = static_cast<unsigned long>(std::llround(gap / res));
        if (ov >= lo.ulRows || ov >= hi.ulRows) { model::doError("A domain lies entirely inside another.", model::errorCodes::kLevelModelStop); return nullptr; }
        off[order[k]] = off[order[k - 1]] + lo.ulRows - ov;
        split[k] = off[order[k]] + ov / 2;
        total = off[order[k]] + hi.ulRows;
    }
    split[n] = total;
    std::unique_ptr<CDomainCartesian> m(new CDomainCartesian());
    m->ulCols = first.ulCols; m->ulRows = total; m->dCellResolution = res; m->dRealOffsetX = first.dRealOffsetX; m->dRealOffsetY = first.dRealOffsetY;
    m->sSourceDir = first.sSourceDir; m->sTargetDir = first.sTargetDir;
    m->dCellStates.assign(4 * m->getCellCount(), 0.0); m->dBedElevations.assign(m->getCellCount(), 0.0); m->dManningValues.assign(m->getCellCount(), 0.0);
    for (size_t k = 0; k < n; ++k) {
        CDomainCartesian& p = *parts[order[k]];
        const unsigned long o = off[order[k]];
        for (unsigned long y = 0; y < p.ulRows; ++y) {
            const unsigned long my = o + y;
            if (my < split[k] || my >= split[k + 1]) continue;           // the neighbour is authoritative there
            std::copy(p.dCellStates.begin() + 4 * y * p.ulCols, p.dCellStates.begin() + 4 * (y + 1) * p.ulCols, m->dCellStates.begin() + 4 * my * m->ulCols);
            std::copy(p.dBedElevations.begin() + y * p.ulCols, p.dBedElevations.begin() + (y + 1) * p.ulCols, m->dBedElevations.begin() + my * m->ulCols);
            std::copy(p.dManningValues.begin() + y * p.ulCols, p.dManningValues.begin() + (y + 1) * p.ulCols, m->dManningValues.begin() + my * m->ulCols);
        }
        // boundaries: cell maps move with their part and keep only the cells the part is authoritative for; domain-wide
        // (atmospheric / gridded) series apply to the merged domain once
        for (auto& b : p.boundaryMap.boundaries) {
            if (auto* cell = dynamic_cast<CBoundaryCell*>(b.get())) {
                std::vector<std::pair<unsigned int, unsigned int>> kept;
                for (const auto& r : cell->relations) { const unsigned long my = o + r.second; if (my >= split[k] && my < split[k + 1]) kept.emplace_back(r.first, static_cast<unsigned int>(my)); }
                if (kept.size() != cell->relations.size() && cell->bDischargeIsTotal)
                    model::doError("Boundary '" + cell->getName() + "': cells in another domain's half of an overlap were dropped.", model::errorCodes::kLevelInformation);
                cell->relations = kept;
                m->boundaryMap.boundaries.push_back(std::move(b));
            } else if (!m->boundaryMap.getBoundaryByName(b->getName())) {
                if (k > 0) model::doError("Boundary '" + b->getName() + "' now applies to the whole merged domain.", model::errorCodes::kLevelInformation);
                m->boundaryMap.boundaries.push_back(std::move(b));
            }
        }
        p.boundaryMap.boundaries.clear();
    }
    return m.release();
}

bool CDomainCartesian::writeCroppedOutputs(double dTime, const CDomainCartesian& merged, unsigned long o, CScheme* pScheme,
                                           std::map<unsigned char, std::vector<double>>& cache) const {
    bool ok = true;
    for (const auto& out : outputs) {
        const unsigned char code = getDataValueCode(out.sValue);
        auto it = cache.find(code);
        if (it == cache.end()) {
            std::vector<double> band;
            if (!(pScheme && pScheme->deriveRaster(code, band))) {
                band.resize(merged.getCellCount());
                for (unsigned long y = 0; y < merged.ulRows; ++y) for (unsigned long x = 0; x < merged.ulCols; ++x) {
                    const unsigned long i = merged.getCellID(x, y);
                    band[(merged.ulRows - 1 - y) * merged.ulCols + x] = deriveOutput(code, &merged.dCellStates[4 * i], merged.dBedElevations[i], merged.dCellResolution, -9999.0);
                }
            }
            it = cache.emplace(code, std::move(band)).first;
        }
        std::string file = out.sTarget; const size_t pos = file.find("%t");
        char tbuf[64]; snprintf(tbuf, sizeof(tbuf), "%g", std::floor(dTime * 100.0) / 100.0);
        if (pos != std::string::npos) file.replace(pos, 2, tbuf);
        // north-first band: this part's northern edge is merged row o + ulRows - 1
        const double* start = it->second.data() + (merged.ulRows - (o + ulRows)) * merged.ulCols;
        ok = CRasterDataset::writeRaster(out.sFormat, sTargetDir + file, ulCols, ulRows, dRealOffsetX, dRealOffsetY, dCellResolution, start) && ok;
    }
    return ok;
}

// ---------------------------------------------------------------------------------------------
// Raster writers
// ---------------------------------------------------------------------------------------------
namespace {
template <class T> void put(std::vector<unsigned char>& b, T v) { const unsigned char* p = reinterpret_cast<const unsigned char*>(&v); b.insert(b.end(), p, p + sizeof(T)); }

// Uncompressed single-band Float64 GeoTIFF, little endian, one strip per row; classic TIFF below 4 GB, BigTIFF above
// (what GDAL's GTiff driver produces with its defaults and BIGTIFF=IF_NEEDED).
bool writeGeoTIFF(const std::string& path, unsigned long cols, unsigned long rows, double ulx, double uly, double res, const double* data) {
    const uint64_t row_bytes = static_cast<uint64_t>(cols) * 8, image_bytes = row_bytes * rows;
    const bool big = image_bytes + 65536 + 16ull * rows >= 0xFFFFFFF0ull;
    struct Entry { uint16_t tag, type; uint64_t count; std::vector<unsigned char> payload; };
    std::vector<Entry> ifd;
    auto add_short = [&](uint16_t tag, uint16_t v) { Entry e{tag, 3, 1, {}}; put<uint16_t>(e.payload, v); ifd.push_back(e); };
    auto add_long = [&](uint16_t tag, uint64_t v) { Entry e{tag, static_cast<uint16_t>(big ? 16 : 4), 1, {}}; if (big) put<uint64_t>(e.payload, v); else put<uint32_t>(e.payload, static_cast<uint32_t>(v)); ifd.push_back(e); };
    auto add_doubles = [&](uint16_t tag, std::initializer_list<double> v) { Entry e{tag, 12, v.size(), {}}; for (double d : v) put<double>(e.payload, d); ifd.push_back(e); };
    auto add_shorts = [&](uint16_t tag, std::initializer_list<uint16_t> v) { Entry e{tag, 3, v.size(), {}}; for (uint16_t d : v) put<uint16_t>(e.payload, d); ifd.push_back(e); };
    auto add_ascii = [&](uint16_t tag, const std::string& t) { Entry e{tag, 2, t.size() + 1, {}}; e.payload.assign(t.begin(), t.end()); e.payload.push_back(0); ifd.push_back(e); };
    const uint64_t header = big ? 16 : 8;
    const uint64_t data_off = header;                                   // pixel data first, IFD after it
    add_long(256, cols); add_long(257, rows); add_short(258, 64); add_short(259, 1); add_short(262, 1);
    { Entry e{273, static_cast<uint16_t>(big ? 16 : 4), rows, {}};
      for (unsigned long r = 0; r < rows; ++r) { if (big) put<uint64_t>(e.payload, data_off + r * row_bytes); else put<uint32_t>(e.payload, static_cast<uint32_t>(data_off + r * row_bytes)); }
      ifd.push_back(e); }
    add_short(277, 1); add_long(278, 1);
    { Entry e{279, static_cast<uint16_t>(big ? 16 : 4), rows, {}};
      for (unsigned long r = 0; r < rows; ++r) { if (big) put<uint64_t>(e.payload, row_bytes); else put<uint32_t>(e.payload, static_cast<uint32_t>(row_bytes)); }
      ifd.push_back(e); }
    add_short(284, 1); add_short(339, 3);
    add_doubles(33550, {res, res, 0.0});                                // ModelPixelScaleTag
    add_doubles(33922, {0.0, 0.0, 0.0, ulx, uly, 0.0});                 // ModelTiepointTag: pixel (0,0) = upper-left corner
    add_shorts(34735, {1, 1, 0, 1, 1025, 0, 1, 1});                     // GeoKeyDirectory: GTRasterTypeGeoKey = RasterPixelIsArea
    add_ascii(42113, "-9999");                                          // GDAL_NODATA
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) return false;
    std::vector<unsigned char> head;
    head.push_back('I'); head.push_back('I');
    const uint64_t ifd_off = (data_off + image_bytes + 7) / 8 * 8;
    if (big) { put<uint16_t>(head, 43); put<uint16_t>(head, 8); put<uint16_t>(head, 0); put<uint64_t>(head, ifd_off); }
    else { put<uint16_t>(head, 42); put<uint32_t>(head, static_cast<uint32_t>(ifd_off)); }
    bool ok = fwrite(head.data(), 1, head.size(), f) == head.size();
    ok = ok && fwrite(data, 1, image_bytes, f) == image_bytes;
    for (uint64_t p = data_off + image_bytes; p < ifd_off && ok; ++p) ok = fputc(0, f) != EOF;
    // IFD: entries sorted by tag (they are added in ascending order), out-of-line payloads behind the table
    const uint64_t entry_size = big ? 20 : 12, inline_cap = big ? 8 : 4;
    const uint64_t table_bytes = (big ? 8 : 2) + entry_size * ifd.size() + (big ? 8 : 4);
    uint64_t extra_off = ifd_off + table_bytes;
    std::vector<unsigned char> table, extra;
    if (big) put<uint64_t>(table, ifd.size()); else put<uint16_t>(table, static_cast<uint16_t>(ifd.size()));
    for (auto& e : ifd) {
        put<uint16_t>(table, e.tag); put<uint16_t>(table, e.type);
        if (big) put<uint64_t>(table, e.count); else put<uint32_t>(table, static_cast<uint32_t>(e.count));
        if (e.payload.size() <= inline_cap) {
            std::vector<unsigned char> v = e.payload; v.resize(inline_cap, 0);
            table.insert(table.end(), v.begin(), v.end());
        } else {
            if (big) put<uint64_t>(table, extra_off + extra.size()); else put<uint32_t>(table, static_cast<uint32_t>(extra_off + extra.size()));
            extra.insert(extra.end(), e.payload.begin(), e.payload.end());
            while (extra.size() % 8) extra.push_back(0);
        }
    }
    if (big) put<uint64_t>(table, 0); else put<uint32_t>(table, 0);     // no further IFD
    ok = ok && fwrite(table.data(), 1, table.size(), f) == table.size();
    ok = ok && (extra.empty() || fwrite(extra.data(), 1, extra.size(), f) == extra.size());
    return fclose(f) == 0 && ok;
}


// ERDAS IMAGINE (.img / HFA) writer: one Float64 layer in uncompressed 64 x 64 blocks, no-data value, map info.
// The node tree and the binary layout of every node follow what GDAL's HFA driver writes (the reference's own test
// DEM is such a file: root -> IMGFormatInfo, Layer_1 -> RasterDMS, Ehfa_Layer, Eimg_NonInitializedValue, Map_Info); the
// MIF dictionary carries the definitions of exactly those object types.  HFA offsets are 32 bit: rasters of 2 GB and
// more are refused (GDAL spills them into a separate .ige file) and the caller falls back to GeoTIFF.
bool writeHFA(const std::string& path, unsigned long cols, unsigned long rows, double ulx, double uly, double res, const double* data) {
    const uint32_t B = 64;
    const uint64_t nbx = (cols + B - 1) / B, nby = (rows + B - 1) / B, nblocks = nbx * nby, block_bytes = uint64_t(B) * B * 8;
    if (nblocks * block_bytes >= 0x7ff00000ull) return false;
    static const char* kDictionary =
        "{1:lversion,1:LfreeList,1:LrootEntryPtr,1:sentryHeaderLength,1:LdictionaryPtr,}Ehfa_File,"
        "{1:Lnext,1:Lprev,1:Lparent,1:Lchild,1:Ldata,1:ldataSize,64:cname,32:ctype,1:tmodTime,}Ehfa_Entry,"
        "{16:clabel,1:LheaderPtr,}Ehfa_HeaderTag,"
        "{1:lwidth,1:lheight,1:e3:thematic,athematic,fft of real-valued data,layerType,"
        "1:e13:u1,u2,u4,u8,s8,u16,s16,u32,s32,f32,f64,c64,c128,pixelType,1:lblockWidth,1:lblockHeight,}Eimg_Layer,"
        "{1:e2:raster,vector,type,1:LdictionaryPtr,}Ehfa_Layer,"
        "{1:LspaceUsedForRasterData,}ImgFormatInfo831,"
        "{1:sfileCode,1:Loffset,1:lsize,1:e2:false,true,logvalid,1:e2:no compression,ESRI GRID compression,compressionType,}Edms_VirtualBlockInfo,"
        "{1:lmin,1:lmax,}Edms_FreeIDList,"
        "{1:lnumvirtualblocks,1:lnumobjectsperblock,1:lnextobjectnum,1:e2:no compression,RLC compression,compressionType,"
        "0:poEdms_VirtualBlockInfo,blockinfo,0:poEdms_FreeIDList,freelist,1:tmodTime,}Edms_State,"
        "{1:*bvalueBD,}Eimg_NonInitializedValue,"
        "{1:dx,1:dy,}Eprj_Coordinate,{1:dwidth,1:dheight,}Eprj_Size,"
        "{0:pcproName,1:*oEprj_Coordinate,upperLeftCenter,1:*oEprj_Coordinate,lowerRightCenter,1:*oEprj_Size,pixelSize,0:pcunits,}Eprj_MapInfo,.";
    std::vector<unsigned char> f;
    auto at = [&]() { return static_cast<uint32_t>(f.size()); };
    auto patch32 = [&](size_t o, uint32_t v) { memcpy(&f[o], &v, 4); };
    const char tag[16] = "EHFA_HEADER_TAG";
    f.insert(f.end(), tag, tag + 16);
    put<uint32_t>(f, 20);                                               // Ehfa_HeaderTag.headerPtr
    put<int32_t>(f, 1); put<uint32_t>(f, 0); const size_t rootPtrAt = f.size(); put<uint32_t>(f, 0); put<int16_t>(f, 128); put<uint32_t>(f, 38);   // Ehfa_File
    f.insert(f.end(), kDictionary, kDictionary + strlen(kDictionary) + 1);
    // entries: 128 bytes {next, prev, parent, child, data, dataSize, name[64], type[32], modTime, pad}
    struct Node { uint32_t at, next, prev, parent, child, data; int32_t size; };
    auto entry = [&](const char* name, const char* type) {
        Node n{at(), 0, 0, 0, 0, 0, 0};
        f.resize(f.size() + 128, 0);
        strncpy(reinterpret_cast<char*>(&f[n.at + 24]), name, 63);
        strncpy(reinterpret_cast<char*>(&f[n.at + 88]), type, 31);
        return n;
    };
    auto commit = [&](const Node& n) { const uint32_t v[5] = {n.next, n.prev, n.parent, n.child, n.data}; memcpy(&f[n.at], v, 20); memcpy(&f[n.at + 20], &n.size, 4); };
    Node root = entry("root", "root");
    patch32(rootPtrAt, root.at);
    Node info = entry("IMGFormatInfo", "ImgFormatInfo831");
    info.parent = root.at; info.data = at(); info.size = 4;
    put<uint32_t>(f, static_cast<uint32_t>(nblocks * block_bytes));
    Node layer = entry("Layer_1", "Eimg_Layer");
    layer.parent = root.at; layer.prev = info.at; info.next = layer.at; root.child = info.at;
    layer.data = at(); layer.size = 20;
    put<int32_t>(f, static_cast<int32_t>(cols)); put<int32_t>(f, static_cast<int32_t>(rows)); put<uint16_t>(f, 1 /* athematic */);
    put<uint16_t>(f, 10 /* f64 */); put<int32_t>(f, B); put<int32_t>(f, B);
    Node dms = entry("RasterDMS", "Edms_State");
    dms.parent = layer.at; layer.child = dms.at; dms.data = at(); dms.size = static_cast<int32_t>(14 + 8 + 14 * nblocks + 16);
    put<int32_t>(f, static_cast<int32_t>(nblocks)); put<int32_t>(f, B * B); put<int32_t>(f, static_cast<int32_t>(nblocks * B * B)); put<uint16_t>(f, 0);
    put<uint32_t>(f, static_cast<uint32_t>(nblocks)); put<uint32_t>(f, at() + 4);       // blockinfo: count, pointer to the array that follows
    const size_t tableAt = f.size();
    f.resize(f.size() + 14 * nblocks, 0);
    f.resize(f.size() + 16, 0);                                          // freelist {0, 0}, modTime, pad
    Node ehl = entry("Ehfa_Layer", "Ehfa_Layer");
    ehl.parent = layer.at; ehl.prev = dms.at; dms.next = ehl.at; ehl.data = at(); ehl.size = 6;
    put<uint16_t>(f, 0 /* raster */); put<uint32_t>(f, at() + 4);
    const char* layerDict = "{4096:ddata,}RasterDMS,.";
    f.insert(f.end(), layerDict, layerDict + strlen(layerDict) + 1);
    Node nd = entry("Eimg_NonInitializedValue", "Eimg_NonInitializedValue");
    nd.parent = layer.at; nd.prev = ehl.at; ehl.next = nd.at; nd.data = at(); nd.size = 28;
    put<uint32_t>(f, 1); put<uint32_t>(f, at() + 4); put<int32_t>(f, 1); put<int32_t>(f, 1); put<uint16_t>(f, 10); put<uint16_t>(f, 0); put<double>(f, -9999.0);
    Node map = entry("Map_Info", "Eprj_MapInfo");
    map.parent = layer.at; map.prev = nd.at; nd.next = map.at; map.data = at();
    const char* pro = "Unknown"; const char* units = "meters";
    put<uint32_t>(f, static_cast<uint32_t>(strlen(pro) + 1)); put<uint32_t>(f, at() + 4); f.insert(f.end(), pro, pro + strlen(pro) + 1);
    put<uint32_t>(f, 1); put<uint32_t>(f, at() + 4); put<double>(f, ulx + 0.5 * res); put<double>(f, uly - 0.5 * res);                               // upperLeftCenter
    put<uint32_t>(f, 1); put<uint32_t>(f, at() + 4); put<double>(f, ulx + res * cols - 0.5 * res); put<double>(f, uly - res * rows + 0.5 * res);     // lowerRightCenter
    put<uint32_t>(f, 1); put<uint32_t>(f, at() + 4); put<double>(f, res); put<double>(f, res);                                                       // pixelSize
    put<uint32_t>(f, static_cast<uint32_t>(strlen(units) + 1)); put<uint32_t>(f, at() + 4); f.insert(f.end(), units, units + strlen(units) + 1);
    map.size = static_cast<int32_t>(at() - map.data);
    for (const Node& n : {root, info, layer, dms, ehl, nd, map}) commit(n);
    // block table, then the blocks themselves (north-west block first, rows of a block north first, edge blocks padded)
    const uint32_t first = at();
    for (uint64_t b = 0; b < nblocks; ++b) {
        unsigned char* rec = &f[tableAt + 14 * b];
        const uint16_t zero = 0, one = 1; const uint32_t off = static_cast<uint32_t>(first + b * block_bytes); const int32_t size = static_cast<int32_t>(block_bytes);
        memcpy(rec, &zero, 2); memcpy(rec + 2, &off, 4); memcpy(rec + 6, &size, 4); memcpy(rec + 10, &one, 2); memcpy(rec + 12, &zero, 2);
    }
    FILE* out = fopen(path.c_str(), "wb");
    if (!out) return false;
    bool ok = fwrite(f.data(), 1, f.size(), out) == f.size();
    std::vector<double> blk(B * B);
    for (uint64_t by = 0; by < nby && ok; ++by) for (uint64_t bx = 0; bx < nbx && ok; ++bx) {
        for (uint32_t y = 0; y < B; ++y) for (uint32_t x = 0; x < B; ++x) {
            const uint64_t gy = by * B + y, gx = bx * B + x;
            blk[y * B + x] = (gy < rows && gx < cols) ? data[gy * cols + gx] : -9999.0;
        }
        ok = fwrite(blk.data(), sizeof(double), blk.size(), out) == blk.size();
    }
    return fclose(out) == 0 && ok;
}

bool writeENVI(const std::string& path, unsigned long cols, unsigned long rows, double ulx, double uly, double res, const double* data) {
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) return false;
    const size_t n = static_cast<size_t>(cols) * rows;
    bool ok = fwrite(data, sizeof(double), n, f) == n;
    ok = fclose(f) == 0 && ok;
    const size_t dot = path.find_last_of('.'), slash = path.find_last_of('/');
    const std::string hdr = ((dot == std::string::npos || (slash != std::string::npos && dot < slash)) ? path : path.substr(0, dot)) + ".hdr";
    f = fopen(hdr.c_str(), "w");
    if (!f) return false;
    fprintf(f, "ENVI\ndescription = {%s}\nsamples = %lu\nlines   = %lu\nbands   = 1\nheader offset = 0\nfile type = ENVI Standard\n"
               "data type = 5\ninterleave = bsq\nbyte order = 0\nmap info = {Arbitrary, 1, 1, %.10g, %.10g, %.10g, %.10g}\n"
               "data ignore value = -9999\nband names = {Band 1}\n", path.c_str(), cols, rows, ulx, uly, res, res);
    return fclose(f) == 0 && ok;
}
}  // namespace

bool CRasterDataset::writeRaster(const std::string& sFormat, const std::string& sFilename, unsigned long cols, unsigned long rows, double offX,
                                 double offY, double res, const double* northFirst, std::string* pWritten) {
    const double ulx = offX, uly = offY + res * rows;                   // top-left instead of bottom-left, CRasterDataset.cpp:166
    std::string file = sFilename;
    bool ok;
    if (sFormat == "AAIGrid") {
        SRaster r; r.cols = cols; r.rows = rows; r.cellsize = res; r.xll = offX; r.yll = offY; r.nodata = -9999.0;
        r.values.resize(static_cast<size_t>(cols) * rows);
        for (unsigned long y = 0; y < rows; ++y) std::copy(northFirst + (rows - 1 - y) * cols, northFirst + (rows - y) * cols, r.values.begin() + y * cols);
        ok = r.write(file);
    } else if (sFormat == "ENVI") {
        ok = writeENVI(file, cols, rows, ulx, uly, res, northFirst);
    } else if (sFormat == "HFA" && writeHFA(file, cols, rows, ulx, uly, res, northFirst)) {
        ok = true;
    } else {
        if (sFormat != "GTiff") {
            model::doError("GDAL format driver '" + sFormat + "' is not available in this build: writing GeoTIFF instead.", model::errorCodes::kLevelWarning);
            const size_t dot = file.find_last_of('.'), slash = file.find_last_of('/');
            file = ((dot == std::string::npos || (slash != std::string::npos && dot < slash)) ? file : file.substr(0, dot)) + ".tif";
        }
        ok = writeGeoTIFF(file, cols, rows, ulx, uly, res, northFirst);
    }
    if (!ok) model::doError("Could not create output raster file.", model::errorCodes::kLevelWarning);     // CRasterDataset.cpp:155-161
    if (pWritten) *pWritten = file;
    return ok;
}

// ---------------------------------------------------------------------------------------------
// Schemes
// ---------------------------------------------------------------------------------------------
CScheme* CScheme::createFromConfig(const XMLElement* pXScheme) {       // src/Schemes/CScheme.cpp:141-175
    const std::string name = Util::toLowercase(pXScheme ? pXScheme->Attribute("name") : nullptr);
    CScheme* s = nullptr;
    if (name == "godunov") s = new CSchemeGodunov();
    else if (name == "muscl-hancock") s = new CSchemeMUSCLHancock();
    else if (name == "inertial") s = new CSchemeInertial();
    else { model::doError("Unsupported scheme specified for the domain.", model::errorCodes::kLevelWarning); return nullptr; }
    s->setupFromConfig(pXScheme);
    return s;
}
CScheme::~CScheme() { cleanupSimulation(); }

void CScheme::setupFromConfig(const XMLElement* pXScheme) {            // CScheme.cpp:79-109, CSchemeGodunov.cpp:128-334
    for (const XMLElement* p = pXScheme->FirstChildElement("parameter"); p; p = p->NextSiblingElement("parameter")) {
        const std::string key = Util::toLowercase(p->Attribute("name")), val = Util::toLowercase(p->Attribute("value"));
        if (key == "courantnumber") { if (isValidFloat(val.c_str())) setCourantNumber(atof(val.c_str())); else model::doError("Invalid Courant number given.", model::errorCodes::kLevelWarning); }
        else if (key == "drythreshold") { if (isValidFloat(val.c_str())) setDryThreshold(atof(val.c_str())); else model::doError("Invalid dry threshold depth given.", model::errorCodes::kLevelWarning); }
        else if (key == "timestepmode") {
            if (val == "auto" || val == "cfl") setTimestepMode(true); else if (val == "fixed") setTimestepMode(false);
            else model::doError("Invalid timestep mode given.", model::errorCodes::kLevelWarning);
        }
        else if (key == "timestepinitial" || key == "timestepfixed") { if (isValidFloat(val.c_str())) setTimestep(atof(val.c_str())); else model::doError("Invalid initial/fixed timestep given.", model::errorCodes::kLevelWarning); }
        else if (key == "frictioneffects") { if (val == "yes") setFrictionStatus(true); else if (val == "no") setFrictionStatus(false); else model::doError("Invalid friction state given.", model::errorCodes::kLevelWarning); }
        else if (key == "queuesize" || key == "queueinitialsize" || key == "queuefixedsize") { if (atoi(val.c_str()) > 0) setQueueSize(atoi(val.c_str())); else model::doError("Invalid queue size given.", model::errorCodes::kLevelWarning); }
        else if (key == "queuemode") {                                              // CScheme.cpp:79-95
            if (val == "auto") setQueueMode(true); else if (val == "fixed") setQueueMode(false);
            else model::doError("Invalid queue mode given.", model::errorCodes::kLevelWarning);
        }
        else if (key == "riemannsolver") { if (val != "hllc") model::doError("Invalid Riemann solver given.", model::errorCodes::kLevelWarning); }
        else if (key == "localcachelevel") {
            // The cache strategy is the executor's business -- except the one observable difference between the two Godunov
            // kernels: gts_cacheEnabled writes nothing when the timestep is <= 0 (CLSchemeGodunov.clc:477-478; selected by
            // src/Schemes/CSchemeGodunov.cpp:296-304, 966-968)
            if (ucSchemeType == model::schemeTypes::kGodunov && (val == "maximum" || val == "max" || val == "enabled")) uiQuirks |= HP_QUIRK_GODUNOV_DT0_KEEP;
            else if (val == "none" || val == "no") uiQuirks &= ~static_cast<uint32_t>(HP_QUIRK_GODUNOV_DT0_KEEP);
        }
        else if (key == "groupsize" || key == "cachedgroupsize" || key == "noncachedgroupsize" || key == "localcacheconstraints" ||
                 key == "timestepreductiondivisions" || key == "contiguousextrapolationdata") { /* launch geometry and cache strategy are the executor's business */ }
        else model::doError("Unrecognised parameter: " + key, model::errorCodes::kLevelWarning);
    }
}

// f(strip, index) for every strip; on several devices each call runs on its own host thread, because the library's
// collective entry points (attach_comm, iterate, update_timestep, prepare_graphs) contain NCCL calls that every rank
// must issue concurrently -- the reference's one worker thread per scheme (CSchemeGodunov.cpp:1116-1141).
template <class F> bool CScheme::forStrips(F f) {
    if (strips.size() <= 1) { if (!strips.empty()) f(strips[0], 0); return !model::forceAbort; }
    std::vector<std::thread> workers;
    for (size_t i = 1; i < strips.size(); ++i) workers.emplace_back([&, i]() { f(strips[i], i); });
    f(strips[0], 0);
    for (auto& w : workers) w.join();
    return !model::forceAbort;
}

bool CScheme::prepareAll(CExecutorControlCUDA* pExec, CDomainCartesian* pDom, unsigned char ucPrecision, double dSimulationLength,
                         const std::vector<int>& devices) {
    pExecutor = pExec; pDomain = pDom; ucFloatPrecision = ucPrecision;
    const unsigned long ulRows = pDom->getRows();
    const unsigned long ulHalo = ucSchemeType == model::schemeTypes::kMUSCLHancock ? 2 : 1;
    size_t n = devices.size() > 1 ? devices.size() : 1;
    if (n > 1 && ulRows / n < 2 * ulHalo + 1) { model::doError("The domain is too short to be split into that many row strips; running on one device.", model::errorCodes::kLevelWarning); n = 1; }
    strips.assign(n, SStrip());
    if (const char* e = getenv("HIPIMS_STRIP_EXCHANGE")) bPeerExchange = std::string(e) == "peer";
    bPeersAttached = false;
    // rows are dealt out as evenly as possible, the first (rows % n) strips get one more (hipims_ocl_b200/strips.py)
    const unsigned long ulBase = ulRows / n, ulExtra = ulRows % n;
    for (size_t i = 0; i < n; ++i) {
        SStrip& st = strips[i];
        st.ulOwnRows = ulBase + (i < ulExtra ? 1 : 0);
        st.ulOwnFirst = i * ulBase + std::min<unsigned long>(i, ulExtra);
        const unsigned long ulHaloS = i > 0 ? ulHalo : 0, ulHaloN = i + 1 < n ? ulHalo : 0;
        st.ulFirstRow = st.ulOwnFirst - ulHaloS; st.ulRows = st.ulOwnRows + ulHaloS + ulHaloN;
        if (i == 0) { st.pExec = pExec->getDevice(); st.bOwnsExecutor = false; }
        else if (hp_executor_create(devices[i], nullptr, &st.pExec) < 0) {
            model::doError(std::string("Could not open a device for a row strip: ") + hp_last_error(), model::errorCodes::kLevelModelStop); cleanupSimulation(); return false;
        } else st.bOwnsExecutor = true;
        hp_scheme_config c{};
        c.struct_size = sizeof(c); c.scheme = ucSchemeType; c.real_bytes = ucPrecision == model::floatPrecision::kSingle ? 4 : 8;
        // Strips of ONE process that exchange through NCCL launch directly: capturing NCCL operations into CUDA graphs from
        // several threads of the same process fails inside NCCL (ncclGroupEnd: internal error, NCCL 2.28, measured); with
        // one process per GPU -- bench.py, tools/multigpu_check.py -- the captured path is used.  Strips that exchange over
        // peer memory (HIPIMS_STRIP_EXCHANGE=peer) have no NCCL call in the loop and replay graphs like a single device.
        c.quirks = uiQuirks; c.options = uiOptions | ((n > 1 && !bPeerExchange) ? HP_OPT_NO_GRAPH : 0u);
        c.dynamic_timestep = bDynamicTimestep ? 1 : 0; c.friction = bFrictionEffects ? 1 : 0;
        c.cols = pDom->getCols(); c.rows = st.ulRows; c.global_rows = ulRows; c.row_offset = st.ulOwnFirst;
        c.halo_south = static_cast<uint32_t>(ulHaloS); c.halo_north = static_cast<uint32_t>(ulHaloN);
        c.delta = pDom->getCellResolution(); c.courant = dCourantNumber; c.dry_threshold = dThresholdVerySmall; c.end_time = dSimulationLength;
        c.fixed_timestep = dTimestep; c.initial_timestep = dTimestep;
        if (hp_scheme_create(st.pExec, &c, &st.pHandle) < 0) {
            model::doError(std::string("Could not prepare the scheme: ") + hp_last_error(), model::errorCodes::kLevelModelStop); cleanupSimulation(); return false;
        }
    }
    pScheme = strips[0].pHandle;
    if (n > 1 && !bPeerExchange) {
        unsigned char id[HP_COMM_ID_BYTES];
        if (hp_comm_unique_id(id) < 0) { model::doError(std::string("NCCL: ") + hp_last_error(), model::errorCodes::kLevelModelStop); cleanupSimulation(); return false; }
        forStrips([&](SStrip& st, size_t i) {
            if (hp_scheme_attach_comm(st.pHandle, id, static_cast<int>(i), static_cast<int>(n)) < 0)
                model::doError(std::string("Could not connect the row strips: ") + hp_last_error(), model::errorCodes::kLevelModelStop);
        });
        if (model::forceAbort) { cleanupSimulation(); return false; }
    }
    // every strip gets every boundary with GLOBAL cell numbers; the library keeps the cells a strip holds
    for (auto& st : strips) pDom->getBoundaries()->prepareBoundaries(st.pHandle, pDom, dSimulationLength);
    dCurrentTimestep = dTimestep;
    return true;
}

void CScheme::prepareSimulation() { prepareSimulationState(); }
void CScheme::prepareSimulationState() {   // CSchemeGodunov.cpp:1053-1071
    if (!pScheme) return;
    const unsigned long ulCols = pDomain->getCols();
    std::vector<float> st32, bed32, man32;
    if (ucFloatPrecision == model::floatPrecision::kSingle) {
        st32.assign(pDomain->dCellStates.begin(), pDomain->dCellStates.end());
        bed32.assign(pDomain->dBedElevations.begin(), pDomain->dBedElevations.end());
        man32.assign(pDomain->dManningValues.begin(), pDomain->dManningValues.end());
    }
    // each strip: the rows it holds, halo rows included.  One host thread per strip: with peers attached an upload ends in
    // a barrier over all strips, so the uploads must be in flight together
    forStrips([&](SStrip& st, size_t) {
        const size_t off = static_cast<size_t>(st.ulFirstRow) * ulCols;
        if (ucFloatPrecision == model::floatPrecision::kSingle)
            HP_CHECK(hp_scheme_upload_cells(st.pHandle, st32.data() + 4 * off, bed32.data() + off, man32.data() + off), "upload");
        else
            HP_CHECK(hp_scheme_upload_cells(st.pHandle, pDomain->dCellStates.data() + 4 * off, pDomain->dBedElevations.data() + off,
                                            pDomain->dManningValues.data() + off), "upload");
        HP_CHECK(hp_scheme_sync(st.pHandle), "upload");
    });
    if (strips.size() > 1 && bPeerExchange && !bPeersAttached) {
        // row strips over peer memory (include/hipims_cuda.h): every strip describes its buffers, then all of them map their
        // peers' -- a rendezvous, after the first upload
        std::vector<unsigned char> blobs(strips.size() * HP_PEER_BLOB_BYTES);
        for (size_t i = 0; i < strips.size(); ++i) HP_CHECK(hp_scheme_peer_export(strips[i].pHandle, blobs.data() + i * HP_PEER_BLOB_BYTES), "peer export");
        forStrips([&](SStrip& st, size_t i) {
            if (hp_scheme_attach_peers(st.pHandle, static_cast<int>(i), static_cast<int>(strips.size()), blobs.data()) < 0)
                model::doError(std::string("Could not connect the row strips over peer memory: ") + hp_last_error(), model::errorCodes::kLevelModelStop);
        });
        bPeersAttached = !model::forceAbort;
    }
    for (auto& st : strips) HP_CHECK(hp_scheme_set_clock(st.pHandle, 0.0, dTimestep, 0.0), "clock");
    ulCurrentCellsCalculated = 0;
}

void CScheme::runSimulation(double dTarget, double dRealTime) {      // CSchemeGodunov.cpp:1374-1453 + one Threaded_runBatch pass
    if (!pScheme) return;
    if (dTarget <= 0.0) { dTargetTime = dTarget; return; }     // "No target time? Can't run anything yet then" (:1384-1386)
    if (dCurrentTime > dTarget + 1E-5) { model::doError("Simulation has exceeded target time", model::errorCodes::kLevelWarning); return; }   // :1389-1405
    // batch size: aim for a second of wall clock per batch (:1419-1450; single domain, so no rollback budget to respect).
    // Decided once, here, for all strips: they must enqueue the same number of iterations.
    if (bAutomaticQueue && dRealTime > 1E-5) {
        const double dBatchDuration = dRealTime - dBatchStartedTime;
        const unsigned int uiOld = uiQueueAdditionSize;
        if (dBatchDuration > 0.0)
            uiQueueAdditionSize = std::max(1u, std::min(uiBatchRate * 3, static_cast<unsigned int>(std::ceil(1.0 / (dBatchDuration / static_cast<double>(uiQueueAdditionSize))))));
        if (uiQueueAdditionSize > uiOld * 2 && uiQueueAdditionSize > 40) uiQueueAdditionSize = std::min(uiBatchRate * 3, uiOld * 2);   // no silly jumps
        if (uiQueueAdditionSize < 1) uiQueueAdditionSize = 1;
    }
    dBatchStartedTime = dRealTime;
    const bool bNewTarget = dTarget != dTargetTime;
    const bool bUpdate = bNewTarget && dCurrentTimestep <= 0.0 && ucSyncMethod == model::syncMethod::kSyncForecast;   // :1191-1196
    // a target that moved in front of the step already decided: the timestep is overridden at the start of the batch (:1198-1232)
    const bool bOverride = bNewTarget && dCurrentTime + dCurrentTimestep > dTarget + 1E-5 && dCurrentTime < dTarget;
    const double dOverride = dTarget - dCurrentTime;
    dTargetTime = dTarget;
    // "Do we need to run any work?" (:1284-1285)
    const unsigned int uiBatch = (dCurrentTime < dTargetTime) ? uiQueueAdditionSize : 0;
    forStrips([&](SStrip& st, size_t) {
        if (bNewTarget) HP_CHECK(hp_scheme_set_target_time(st.pHandle, dTarget), "target time");
        if (bUpdate) HP_CHECK(hp_scheme_update_timestep(st.pHandle), "timestep update");
        if (bOverride) HP_CHECK(hp_scheme_force_timestep(st.pHandle, dOverride), "timestep override");
        if (uiBatch) HP_CHECK(hp_scheme_iterate(st.pHandle, uiBatch), "iterate");
        HP_CHECK(hp_scheme_sync(st.pHandle), "sync");
    });
    ulCurrentCellsCalculated += static_cast<unsigned long long>(uiBatch) * pDomain->getCellCount();
    readKeyStatistics();
}

void CScheme::readKeyStatistics() {
    hp_scheme_stats st{};
    if (hp_scheme_read_stats(pScheme, &st) < 0) { model::doError(hp_last_error(), model::errorCodes::kLevelModelStop); return; }
    // every strip runs the same time controller on the same all-reduced maximum: their clocks are bit-identical
    for (size_t i = 1; i < strips.size(); ++i) {
        hp_scheme_stats o{};
        if (hp_scheme_read_stats(strips[i].pHandle, &o) < 0 || o.time != st.time || o.timestep != st.timestep || o.batch_successful != st.batch_successful)
            model::doError("The clocks of the row strips have diverged.", model::errorCodes::kLevelModelStop);
    }
    uiBatchRate = st.batch_successful > uiBatchSuccessful ? st.batch_successful - uiBatchSuccessful : 1;    // CSchemeGodunov.cpp:1834
    dCurrentTime = st.time; dCurrentTimestep = st.timestep; dBatchTimesteps = st.batch_timesteps;
    uiBatchSuccessful = st.batch_successful; uiBatchSkipped = st.batch_skipped;
}

void CScheme::readDomainAll() {
    if (!pScheme) return;
    const unsigned long ulCols = pDomain->getCols();
    for (auto& st : strips) {                                    // the OWNED rows of every strip, straight into the domain's arrays
        const size_t off = static_cast<size_t>(st.ulOwnFirst) * ulCols * 4;
        const unsigned long ulLocal = st.ulOwnFirst - st.ulFirstRow;
        if (ucFloatPrecision == model::floatPrecision::kSingle) {
            std::vector<float> tmp(static_cast<size_t>(st.ulOwnRows) * ulCols * 4);
            HP_CHECK(hp_scheme_read_rows(st.pHandle, ulLocal, st.ulOwnRows, tmp.data()), "download");
            std::copy(tmp.begin(), tmp.end(), pDomain->dCellStates.begin() + off);
        } else {
            HP_CHECK(hp_scheme_read_rows(st.pHandle, ulLocal, st.ulOwnRows, pDomain->dCellStates.data() + off), "download");
        }
    }
}
bool CScheme::deriveRaster(unsigned char ucValue, std::vector<double>& northFirst) {
    if (!pScheme) return false;
    northFirst.resize(pDomain->getCellCount());
    const unsigned long ulCols = pDomain->getCols(), ulRows = pDomain->getRows();
    for (auto& st : strips) {        // a strip's owned rows [a, b) are the north-first rows [rows - b, rows - a) of the band
        double* out = northFirst.data() + static_cast<size_t>(ulRows - (st.ulOwnFirst + st.ulOwnRows)) * ulCols;
        if (hp_scheme_derive_raster(st.pHandle, ucValue, -9999.0, out) < 0) { model::doError(hp_last_error(), model::errorCodes::kLevelWarning); return false; }
    }
    return true;
}
bool CScheme::isSimulationSyncReady(double dExpected) const { return !(dExpected - dCurrentTime > 1E-5); }       // CSchemeGodunov.cpp:1568-1612
bool CScheme::isSimulationFailure(double dExpected) const {                                                      // :1523-1555
    // can't exceed the number of buffer cells in forecast mode; in timestep mode this "shouldn't happen"
    if (ucSyncMethod == model::syncMethod::kSyncForecast && uiBatchSuccessful >= uiRollbackLimit && dExpected - dCurrentTime > 1E-5) return true;
    if (ucSyncMethod == model::syncMethod::kSyncTimestep && uiBatchSuccessful > uiRollbackLimit) return true;
    if (dCurrentTime > dExpected + 1E-5) { model::doError("Scheme has exceeded target sync time. Rolling back...", model::errorCodes::kLevelWarning); return true; }
    return false;
}
void CScheme::rollbackSimulation(double dTime, double dTarget) {                                                 // :1474-1518
    if (!pScheme) return;
    prepareSimulationState();
    dCurrentTime = dTime; dTargetTime = dTarget;
    forStrips([&](SStrip& st, size_t) {
        HP_CHECK(hp_scheme_set_clock(st.pHandle, dTime, dCurrentTimestep, 0.0), "rollback clock");
        HP_CHECK(hp_scheme_set_target_time(st.pHandle, dTarget), "rollback target");
        if (bDynamicTimestep) HP_CHECK(hp_scheme_update_timestep(st.pHandle), "rollback timestep");              // tst_Reduce + tst_UpdateTimestep
        HP_CHECK(hp_scheme_reset_counters(st.pHandle), "rollback counters");
        HP_CHECK(hp_scheme_sync(st.pHandle), "rollback");
    });
    readKeyStatistics();
}
double CScheme::proposeSyncPoint(double dTime) const {                                                           // :1758-1790
    const double dLimit = 999999999.0, dSpares = 3.0;                   // CDomain.cpp:45, CDomainManager default spare iterations
    double dProposal = dTime + std::fabs(dTimestep);
    if (dTime > 1E-5 && uiBatchSuccessful > 0)
        dProposal = dTime + std::max(std::fabs(dTimestep), dLimit * (dBatchTimesteps / uiBatchSuccessful) * ((dLimit - dSpares) / dLimit));
    else if (dProposal - dTime < 1E-5) dProposal = dTime + std::fabs(dTimestep);
    return dProposal;
}
void CScheme::forceTimestep(double dt) { for (auto& st : strips) HP_CHECK(hp_scheme_force_timestep(st.pHandle, dt), "force timestep"); }
void CScheme::cleanupSimulation() {
    // schemes first (their NCCL communicators are torn down collectively: one thread per strip), then the extra executors
    if (strips.size() > 1) {
        std::vector<std::thread> workers;
        for (auto& st : strips) workers.emplace_back([&st]() { if (st.pHandle) hp_scheme_destroy(st.pHandle); st.pHandle = nullptr; });
        for (auto& w : workers) w.join();
    } else if (!strips.empty() && strips[0].pHandle) { hp_scheme_destroy(strips[0].pHandle); strips[0].pHandle = nullptr; }
    for (auto& st : strips) if (st.bOwnsExecutor && st.pExec) hp_executor_destroy(st.pExec);
    strips.clear();
    pScheme = nullptr;
}

// ---------------------------------------------------------------------------------------------
// Model
// ---------------------------------------------------------------------------------------------
CModel::~CModel() { pScheme.reset(); pExecutor.reset(); }

bool CModel::loadConfiguration(const std::string& sPath, bool bDeviceless) {
    XMLDocument doc;
    if (!doc.LoadFile(sPath)) { model::doError("Cannot load configuration: " + doc.error, model::errorCodes::kLevelModelStop); return false; }
    const XMLElement* root = doc.RootElement();
    if (!root || std::string(root->Name()) != "configuration") { model::doError("Configuration file has no <configuration> root.", model::errorCodes::kLevelModelStop); return false; }
    const size_t slash = sPath.find_last_of('/');
    const std::string dir = slash == std::string::npos ? "" : sPath.substr(0, slash + 1);
    if (const XMLElement* m = root->FirstChildElement("metadata")) {
        if (m->FirstChildElement("name")) sName = m->FirstChildElement("name")->text;
        if (m->FirstChildElement("description")) sDescription = m->FirstChildElement("description")->text;
    }
    if (!bDeviceless) {
        pExecutor.reset(CExecutorControlCUDA::createFromConfig(root->FirstChildElement("execution")));
        if (!pExecutor) return false;
    }
    const XMLElement* sim = root->FirstChildElement("simulation");
    if (!sim) { model::doError("No <simulation> element.", model::errorCodes::kLevelModelStop); return false; }
    for (const XMLElement* p = sim->FirstChildElement("parameter"); p; p = p->NextSiblingElement("parameter")) {   // CModel.cpp:65-135
        const std::string key = Util::toLowercase(p->Attribute("name")), val = Util::toLowercase(p->Attribute("value"));
        if (key == "duration") { if (isValidFloat(val.c_str())) dSimulationTime = atof(val.c_str()); else model::doError("Invalid simulation length given.", model::errorCodes::kLevelWarning); }
        else if (key == "outputfrequency") { if (isValidFloat(val.c_str())) dOutputFrequency = atof(val.c_str()); else model::doError("Invalid output frequency given.", model::errorCodes::kLevelWarning); }
        else if (key == "floatingpointprecision") {
            if (val == "single") ucFloatPrecision = model::floatPrecision::kSingle; else if (val == "double") ucFloatPrecision = model::floatPrecision::kDouble;
            else model::doError("Invalid float precision given.", model::errorCodes::kLevelWarning);
        }
        else if (key == "realstart") { /* wall-clock labelling only */ }
        else model::doError("Unrecognised parameter: " + key, model::errorCodes::kLevelWarning);
    }
    const XMLElement* set = sim->FirstChildElement("domainSet");
    if (set) {                                                     // src/Domain/CDomainManager.cpp:56-76
        const std::string sSync = Util::toLowercase(set->Attribute("syncMethod"));
        if (sSync == "timestep") ucSyncMethod = model::syncMethod::kSyncTimestep;
        else if (sSync == "forecast" || sSync.empty()) ucSyncMethod = model::syncMethod::kSyncForecast;
        else model::doError("Unrecognised synchronisation method: " + sSync, model::errorCodes::kLevelWarning);
    }
    const XMLElement* dom = set ? set->FirstChildElement("domain") : nullptr;
    if (!dom) { model::doError("No <domain> defined.", model::errorCodes::kLevelModelStop); return false; }
    const XMLElement* pXScheme = dom->FirstChildElement("scheme");
    // <domain deviceNumber="N"> (src/Domain/CDomain.cpp:88, CDomainManager.cpp:140-177): the distinct devices the domains
    // name, south to north in the order they appear, are the GPUs the (merged) domain is spread over as row strips
    std::vector<int> xmlDevices;
    for (const XMLElement* d = dom; d; d = d->NextSiblingElement("domain")) {
        const char* a = d->Attribute("deviceNumber");
        if (!a || !*a) continue;
        const int n = atoi(a);
        if (n >= 1 && std::find(xmlDevices.begin(), xmlDevices.end(), n) == xmlDevices.end()) xmlDevices.push_back(n);
    }
    if (!dom->NextSiblingElement("domain")) {
        if (Util::toLowercase(dom->Attribute("type")) != "cartesian") { model::doError("Unsupported domain type.", model::errorCodes::kLevelModelStop); return false; }
        pDomain.reset(new CDomainCartesian());
        // boundary files are relative to the configuration file
        if (!pDomain->configureDomain(dom, dir)) return false;
    } else {
        // several <domain>s (CDomainManager.cpp:56-282): merged into one, see CDomainCartesian::mergeStacked
        const std::string schemeName = Util::toLowercase(pXScheme ? pXScheme->Attribute("name") : nullptr);
        for (const XMLElement* d = dom; d; d = d->NextSiblingElement("domain")) {
            if (Util::toLowercase(d->Attribute("type")) != "cartesian") { model::doError("Unsupported domain type.", model::errorCodes::kLevelModelStop); return false; }
            parts.emplace_back(new CDomainCartesian());
            if (!parts.back()->configureDomain(d, dir)) return false;
            const XMLElement* sch = d->FirstChildElement("scheme");
            if (Util::toLowercase(sch ? sch->Attribute("name") : nullptr) != schemeName)
                model::doError("Domains name different schemes; the first domain's scheme is used for the merged domain.", model::errorCodes::kLevelWarning);
        }
        pDomain.reset(CDomainCartesian::mergeStacked(parts, partRowOffsets));
        if (!pDomain) return false;
    }
    pScheme.reset(CScheme::createFromConfig(pXScheme));
    if (!pScheme) return false;
    if (bDeviceless) return true;      // host-side parsing only (CPU tests)
    // which devices: --devices / HIPIMS_DEVICES if given, else the configuration's device numbers; numbers beyond the
    // devices present are dropped with a warning (a configuration written for four GPUs still runs on one).  The
    // executor's own device (<executor><parameter name="deviceNumber">) always comes first.
    std::vector<int> numbers = stripDeviceNumbers;
    if (numbers.empty()) if (const char* e = getenv("HIPIMS_DEVICES")) for (const char* p = e; *p;) { numbers.push_back(atoi(p)); while (*p && *p != ',') ++p; if (*p) ++p; }
    if (numbers.empty() && !parts.empty()) numbers = xmlDevices;
    std::vector<int> ordinals;
    for (int n : numbers) {
        if (n < 1 || n > static_cast<int>(pExecutor->getDeviceCount())) { model::doError("Device " + std::to_string(n) + " named by the configuration is not present; its rows go to the devices that are.", model::errorCodes::kLevelWarning); continue; }
        if (std::find(ordinals.begin(), ordinals.end(), n - 1) == ordinals.end()) ordinals.push_back(n - 1);
    }
    if (ordinals.size() > 1) {
        const int first = pExecutor->getDeviceOrdinal();
        auto it = std::find(ordinals.begin(), ordinals.end(), first);
        if (it == ordinals.end()) ordinals.insert(ordinals.begin(), first); else std::rotate(ordinals.begin(), it, it + 1);
    }
    return pScheme->prepareAll(pExecutor.get(), pDomain.get(), ucFloatPrecision, dSimulationTime, ordinals);
}

// ---- the management loop (src/CModel.cpp) ---------------------------------------------------------------------------------
bool CModel::runModel() {                                                      // :217-262
    if (!pScheme || !pScheme->isReady()) return false;
    runModelPrepare();
    runModelMain();
    runModelCleanup();
    return !model::forceAbort;
}

void CModel::runModelPrepare() {                                               // :497-547
    model::forceAbort = false;
    // "Can't have timestep sync if we've only got one domain" (:503-505) -- and a decomposed model IS one domain here
    if (ucSyncMethod == model::syncMethod::kSyncTimestep) ucSyncMethod = model::syncMethod::kSyncForecast;
    pScheme->setSyncMethod(ucSyncMethod);
    pScheme->prepareSimulation();
    pScheme->setRollbackLimit();                                               // no links: not constrained by overlapping
    bSynchronised = true; bAllIdle = true; bRollbackRequired = false; bWaitOnLinks = false;
    dTargetTime = 0.0; dLastSyncTime = -1.0; dLastOutputTime = 0.0; dCurrentTime = 0.0; dEarliestTime = 0.0; dGlobalTimestep = 0.0;
    uiSyncCount = 0; uiOutputCount = 0;
}

void CModel::runModelDomainAssess(bool* bSyncReady, bool* bIdle) {             // :552-692
    bRollbackRequired = false; dEarliestTime = 0.0; bWaitOnLinks = false;
    dEarliestTime = pScheme->getCurrentTime();                                 // the minimum over the local domains
    // either we're not ready to sync, or we were still synced from the last run
    if (!pScheme->isSimulationSyncReady(dTargetTime) || bSynchronised || dLastSyncTime == dEarliestTime) {
        bSyncReady[0] = false;
        if (pScheme->isSimulationFailure(dTargetTime)) bRollbackRequired = true;
    } else {
        bSyncReady[0] = true;
    }
    bIdle[0] = true;                          // CScheme::runSimulation returns once its batch has left the device
    bSynchronised = bSyncReady[0]; bAllIdle = bIdle[0];
    if (bAllIdle && !bWaitOnLinks) {
        dGlobalTimestep = pScheme->getCurrentTimestep() > 0.0 ? pScheme->getCurrentTimestep() : 0.0;
        dCurrentTime = dEarliestTime;
    }
}

void CModel::runModelDomainExchange() {}      // :697-713 importLinkZoneData: the strips exchanged their halo rows on the device

void CModel::runModelUpdateTarget(double dTimeBase) {                          // :718-770
    double dEarliestSyncProposal = dSimulationTime;
    // several domains in forecast mode would take the smallest proposeSyncPoint here (:730-738); one domain runs free
    // until outputs are needed.  Don't exceed an output interval:
    const double dFrequency = dOutputFrequency > 0.0 ? dOutputFrequency : dSimulationTime;
    if (std::floor(dEarliestSyncProposal / dFrequency) > std::floor(dLastSyncTime / dFrequency))
        dEarliestSyncProposal = (std::floor(dLastSyncTime / dFrequency) + 1.0) * dFrequency;
    (void)dTimeBase;
    dTargetTime = dEarliestSyncProposal;
}

void CModel::writeOutputs() {
    if (parts.empty()) {
        pDomain->writeOutputs(dCurrentTime, pScheme.get());                    // derived on the device; no full-state read-back
    } else {                                                                   // every original domain gets its own rasters
        std::map<unsigned char, std::vector<double>> bands;
        for (size_t i = 0; i < parts.size(); ++i) parts[i]->writeCroppedOutputs(dCurrentTime, *pDomain, partRowOffsets[i], pScheme.get(), bands);
    }
    ++uiOutputCount;
}

void CModel::runModelOutputs() {                                               // :859-882
    const double dFrequency = dOutputFrequency > 0.0 ? dOutputFrequency : dSimulationTime;
    if (bRollbackRequired || !bSynchronised || !bAllIdle ||
        !(std::fabs(dCurrentTime - dLastOutputTime - dFrequency) < 1E-5 && dCurrentTime > dLastOutputTime))
        return;
    writeOutputs();
    dLastOutputTime = dCurrentTime;
    pScheme->forceTimeAdvance();
}

void CModel::runModelSync() {                                                  // :775-835
    if (bRollbackRequired || !bSynchronised || !bAllIdle) return;
    // no rollback required, thus the simulation time can now be increased to match the target
    dCurrentTime = dEarliestTime;
    dLastSyncTime = dCurrentTime;
    ++uiSyncCount;
    runModelOutputs();
    runModelUpdateTarget(dCurrentTime);
    // the state goes back to host memory only where a rollback could need it (several domains in forecast mode) or
    // outputs are about to be written from it (:808-817); the rasters here are derived on the device
    const double dFrequency = dOutputFrequency > 0.0 ? dOutputFrequency : dSimulationTime;
    if (std::fabs(dCurrentTime - dLastOutputTime - dFrequency) < 1E-5 && dCurrentTime > dLastOutputTime) pScheme->saveCurrentState();
    runModelDomainExchange();
}

void CModel::runModelSchedule(double dSeconds, bool* bIdle) {                  // :897-943
    // keep running each domain until we're ready for synchronisation
    if (!bSynchronised && bIdle[0]) {
        if (ucSyncMethod == model::syncMethod::kSyncTimestep && dGlobalTimestep > 0.0) pScheme->forceTimestep(dGlobalTimestep);
        pScheme->runSimulation(dTargetTime, dSeconds);
    }
}

void CModel::runModelRollback() {                                              // :964-1021
    if (!bRollbackRequired || model::forceAbort || !bAllIdle) return;
    model::doError("Rollback invoked - code not yet ready", model::errorCodes::kLevelModelStop);     // the reference stops here too (:971-974)
    bRollbackRequired = false; bSynchronised = false;
    runModelUpdateTarget(dLastSyncTime);
    dEarliestTime = dLastSyncTime; dCurrentTime = dLastSyncTime;
    pScheme->rollbackSimulation(dLastSyncTime, dTargetTime);
}

void CModel::runModelCleanup() {                                               // :1027-1035
    pScheme->readDomainAll();                 // final state back in the CDomain arrays (the strips stay until the model goes)
}

void CModel::runModelMain() {                                                  // :1041-1139
    bool bSyncReady[1] = {false}, bIdle[1] = {true};
    const auto tStart = std::chrono::steady_clock::now();
    // even if the user has forced an abort, still wait until the all-idle state is reached
    while ((dCurrentTime < dSimulationTime - 1E-5 && !model::forceAbort) || !bAllIdle) {
        runModelDomainAssess(bSyncReady, bIdle);
        runModelRollback();
        runModelSync();
        if (bRollbackRequired) continue;
        const double dSeconds = bRealTimeQueue ? std::chrono::duration<double>(std::chrono::steady_clock::now() - tStart).count() + 1E-4 : 0.0;
        runModelSchedule(dSeconds, bIdle);
    }
}

// ---------------------------------------------------------------------------------------------
// C facade for the tests (ctypes)
// ---------------------------------------------------------------------------------------------
extern "C" {
void* hph_model_load(const char* path, int deviceless) {
    model::forceAbort = false; model::errorLog.clear();
    CModel* m = new CModel();
    if (!m->loadConfiguration(path, deviceless != 0)) { delete m; return nullptr; }
    return m;
}
int hph_model_run(void* h) { return static_cast<CModel*>(h)->runModel() ? 0 : -1; }
void hph_model_destroy(void* h) { delete static_cast<CModel*>(h); }
void hph_model_info(void* h, unsigned long* cols, unsigned long* rows, double* resolution, double* duration, double* output_frequency,
                    int* precision, int* scheme, unsigned int* boundaries) {
    CModel* m = static_cast<CModel*>(h);
    *cols = m->getDomain()->getCols(); *rows = m->getDomain()->getRows(); *resolution = m->getDomain()->getCellResolution();
    *duration = m->getSimulationLength(); *output_frequency = m->getOutputFrequency(); *precision = m->getFloatPrecision();
    *scheme = m->getScheme()->getSchemeType(); *boundaries = m->getDomain()->getBoundaries()->getBoundaryCount();
}
unsigned int hph_model_parts(void* h, unsigned long* row_offsets, unsigned int capacity) {
    CModel* m = static_cast<CModel*>(h);
    for (unsigned int i = 0; i < m->getPartCount() && i < capacity; ++i) row_offsets[i] = m->getPartRowOffset(i);
    return m->getPartCount();
}
unsigned int hph_model_strips(void* h) { return static_cast<CModel*>(h)->getStripCount(); }
// like hph_model_load, with the (1-based) devices to spread the domain over given by the caller
void* hph_model_load_on(const char* path, const int* devices, int count) {
    model::forceAbort = false; model::errorLog.clear();
    CModel* m = new CModel();
    m->setStripDevices(std::vector<int>(devices, devices + count));
    if (!m->loadConfiguration(path, false)) { delete m; return nullptr; }
    return m;
}
// the management loop's bookkeeping after (or before) a run: synchronisations, output sets written, target, last sync time
void hph_model_loop_state(void* h, unsigned int* syncs, unsigned int* outputs, double* target, double* last_sync, double* current, int* sync_method) {
    CModel* m = static_cast<CModel*>(h);
    *syncs = m->getSyncCount(); *outputs = m->getOutputCount(); *target = m->getTargetTime(); *last_sync = m->getLastSyncTime();
    *current = m->getCurrentTime(); *sync_method = m->getSyncMethod();
}
double hph_model_next_target(void* h, double last_sync) { return static_cast<CModel*>(h)->proposeTargetAfterSyncAt(last_sync); }
void hph_model_set_realtime_queue(void* h, int on) { static_cast<CModel*>(h)->setRealTimeQueue(on != 0); }
// rollback: put the host arrays and clock back on the device, then report the recomputed timestep
double hph_model_rollback(void* h, double time, double target) {
    CScheme* s = static_cast<CModel*>(h)->getScheme();
    s->rollbackSimulation(time, target);
    return s->getCurrentTimestep();
}
double hph_model_propose_sync(void* h, double time) { return static_cast<CModel*>(h)->getScheme()->proposeSyncPoint(time); }
void hph_model_scheme_params(void* h, double* courant, double* dry, double* timestep, int* dynamic, int* friction, unsigned int* queue) {
    CScheme* s = static_cast<CModel*>(h)->getScheme();
    *courant = s->dCourantNumber; *dry = s->dThresholdVerySmall; *timestep = s->dTimestep; *dynamic = s->bDynamicTimestep; *friction = s->bFrictionEffects;
    *queue = s->getBatchSize();
}
const double* hph_model_states(void* h) { return static_cast<CModel*>(h)->getDomain()->dCellStates.data(); }
const double* hph_model_bed(void* h) { return static_cast<CModel*>(h)->getDomain()->dBedElevations.data(); }
const double* hph_model_manning(void* h) { return static_cast<CModel*>(h)->getDomain()->dManningValues.data(); }
void hph_model_clock(void* h, double* time, double* timestep, unsigned int* ok, unsigned int* skipped) {
    CScheme* s = static_cast<CModel*>(h)->getScheme();
    *time = s->getCurrentTime(); *timestep = s->getCurrentTimestep(); *ok = s->getIterationsSuccessful(); *skipped = s->getIterationsSkipped();
}
// boundary i: kind 0 uniform / 1 gridded / 2 cell; returns the series length in doubles
int hph_model_boundary(void* h, unsigned int i, int* kind, int* def_a, int* def_b, double* interval, double* length, const double** series,
                       unsigned int* relations) {
    CBoundaryMap* map = static_cast<CModel*>(h)->getDomain()->getBoundaries();
    if (i >= map->boundaries.size()) return -1;
    CBoundary* b = map->boundaries[i].get();
    *relations = 0; *series = nullptr; *def_b = 0;
    if (auto* u = dynamic_cast<CBoundaryUniform*>(b)) { *kind = 0; *def_a = u->ucValue; *interval = u->dTimeseriesInterval; *length = u->dTimeseriesLength; *series = u->series.data(); return static_cast<int>(u->series.size()); }
    if (auto* c = dynamic_cast<CBoundaryCell*>(b)) { *kind = 2; *def_a = c->ucDepthValue; *def_b = c->ucDischargeValue; *interval = c->dTimeseriesInterval; *length = c->dTimeseriesLength; *series = c->series.data(); *relations = static_cast<unsigned int>(c->relations.size()); return static_cast<int>(c->series.size()); }
    if (auto* g = dynamic_cast<CBoundaryGridded*>(b)) { *kind = 1; *def_a = g->ucValue; *interval = g->dInterval; *length = 0; return 0; }
    return -1;
}
// raster reader probe: fills cols/rows/cellsize/xll/yll and returns a malloc'ed south-first array (free with hph_free)
double* hph_raster_read(const char* path, unsigned long* cols, unsigned long* rows, double* cellsize, double* xll, double* yll) {
    SRaster r;
    if (!r.read(path)) return nullptr;
    *cols = r.cols; *rows = r.rows; *cellsize = r.cellsize; *xll = r.xll; *yll = r.yll;
    double* out = static_cast<double*>(malloc(r.values.size() * sizeof(double)));
    memcpy(out, r.values.data(), r.values.size() * sizeof(double));
    return out;
}
void hph_free(void* p) { free(p); }
int hph_error_count(void) { return static_cast<int>(model::errorLog.size()); }
const char* hph_error(int i) { return (i >= 0 && i < static_cast<int>(model::errorLog.size())) ? model::errorLog[i].c_str() : ""; }
double hph_round(double v, int places) { return Util::round(v, static_cast<unsigned char>(places)); }
int hph_write_raster(const char* format, const char* path, unsigned long cols, unsigned long rows, double off_x, double off_y, double res,
                     const double* north_first, char* written, size_t written_len) {
    std::string out;
    const bool ok = CRasterDataset::writeRaster(format, path, cols, rows, off_x, off_y, res, north_first, &out);
    if (written && written_len) { strncpy(written, out.c_str(), written_len - 1); written[written_len - 1] = 0; }
    return ok ? 0 : -1;
}
double hph_derive_output(const char* value, const double* state4, double bed, double resolution) {
    return CDomainCartesian::deriveOutput(CDomainCartesian::getDataValueCode(value), state4, bed, resolution, -9999.0);
}
}
