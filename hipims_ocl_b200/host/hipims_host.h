// hipims_host.h -- C++ host side above the C ABI (include/hipims_cuda.h), mirroring the
// reference's executor / scheme / domain / boundary surface for the hot path.
//
// Class and method names, XML attributes and error behaviour follow the reference so that the
// parity tests read like its own code would:
//   model::doError                    src/main.cpp:631-652        (levels; ModelStop sets forceAbort)
//   CExecutorControl::createFromConfig src/Base/CExecutorControl.cpp:66-98 (executor name "CUDA")
//   CScheme (abstract)                src/Schemes/CScheme.h:73-162
//   CSchemeGodunov / MUSCLHancock / Inertial   src/Schemes/CScheme*.cpp
//   CDomainCartesian                  src/Domain/Cartesian/CDomainCartesian.cpp, src/Domain/CDomain.cpp
//   CBoundary / CBoundaryCell / Uniform / Gridded / CBoundaryMap   src/Boundaries/*.cpp
//   CModel (minimal driver)           src/CModel.cpp:217, 1041-1139
// The classes talk to the device ONLY through the C ABI.  They do not implement the reference's
// control plane (polling UI, MPI, GDAL); rasters are ESRI ASCII grids.
#pragma once

#include <cstdint>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/hipims_cuda.h"

namespace model {
namespace errorCodes { enum errorCodes { kLevelFatal = 1, kLevelModelStop = 2, kLevelModelContinue = 4, kLevelWarning = 8, kLevelInformation = 16 }; }
namespace floatPrecision { enum floatPrecision { kSingle = 0, kDouble = 1 }; }
namespace schemeTypes { enum schemeTypes { kGodunov = 0, kMUSCLHancock = 1, kInertialSimplification = 2 }; }
namespace syncMethod { enum syncMethod { kSyncTimestep = 0, kSyncForecast = 1 }; }     // src/Schemes/CScheme.h:57-62
extern bool forceAbort;
extern std::vector<std::string> errorLog;       // every doError message, newest last
void doError(const std::string& message, unsigned char level);
}  // namespace model

// ---------------------------------------------------------------------------------------------
// Minimal XML DOM with the tinyxml2 calls the reference's config code uses.
// ---------------------------------------------------------------------------------------------
class XMLElement {
  public:
    const char* Name() const { return name.c_str(); }
    const char* Attribute(const char* key) const;                 // NULL when absent
    const XMLElement* FirstChildElement(const char* tag = nullptr) const;
    const XMLElement* NextSiblingElement(const char* tag = nullptr) const;
    std::string name, text;
    std::vector<std::pair<std::string, std::string>> attributes;
    std::vector<std::unique_ptr<XMLElement>> children;
    const XMLElement* parent = nullptr;
};
class XMLDocument {
  public:
    bool Parse(const std::string& xml);
    bool LoadFile(const std::string& path);
    const XMLElement* RootElement() const { return root.get(); }
    std::string error;
  private:
    std::unique_ptr<XMLElement> root;
};

// src/Datasets/CCSVDataset.*: rows of comma separated cells, first row is a header
class CCSVDataset {
  public:
    explicit CCSVDataset(const std::string& path) : sFilename(path) {}
    bool readFile();
    bool isReady() const { return bReady; }
    std::vector<std::vector<std::string>> rows;
  private:
    std::string sFilename;
    bool bReady = false;
};

// GDAL-free stand-in for CRasterDataset: reads ESRI ASCII grids and ERDAS IMAGINE (.img / HFA) single-band
// rasters (uncompressed and run-length "ESRI GRID" compressed 64x64 blocks -- the reference's test DEM),
// writes ESRI ASCII grids.  Rows are stored south-first like the reference's cell arrays
// (src/Datasets/CRasterDataset.cpp:411 flips on load).
struct SRaster {
    unsigned long cols = 0, rows = 0;
    double xll = 0, yll = 0, cellsize = 1, nodata = -9999;
    std::vector<double> values;
    bool read(const std::string& path);          // by content: "EHFA_HEADER_TAG" -> HFA, else ASCII grid
    bool readASCII(const std::string& path);
    bool readHFA(const std::string& path);
    bool write(const std::string& path) const;
};

// GDAL-free stand-in for the writing half of CRasterDataset (src/Datasets/CRasterDataset.cpp:101-287): one
// Float64 band, no-data -9999, geotransform {xll, res, 0, yll + res * rows, 0, -res} (:163-170).  `format` keeps
// GDAL's driver codes: "GTiff" (uncompressed GeoTIFF, BigTIFF above 4 GB), "HFA" (ERDAS IMAGINE .img as GDAL lays it
// out, below 2 GB), "ENVI" (raw little-endian + .hdr), "AAIGrid" (ESRI ASCII).  Other codes are reported through
// doError(kLevelWarning) like a driver that cannot create files (:129-146) and written as GeoTIFF next to the
// requested name.
class CRasterDataset {
  public:
    // `northFirst`: rows x cols values, row 0 = northern edge (raster order)
    static bool writeRaster(const std::string& sFormat, const std::string& sFilename, unsigned long ulCols, unsigned long ulRows,
                            double dOffsetX, double dOffsetY, double dResolution, const double* northFirst, std::string* pWritten = nullptr);
};

namespace Util {
double round(double value, unsigned char places);   // src/util.cpp:79-93 (negatives go towards -inf)
std::string toLowercase(const char* s);
}

class CDomainCartesian;
class CScheme;

// ---------------------------------------------------------------------------------------------
// Executor: replaces CExecutorControlOpenCL + COCLDevice.
// ---------------------------------------------------------------------------------------------
class CExecutorControlCUDA {
  public:
    static CExecutorControlCUDA* createFromConfig(const XMLElement* pXExecution);   // <executor name="CUDA"|"OpenCL">
    ~CExecutorControlCUDA();
    bool setupFromConfig(const XMLElement* pXExecutor);      // parameter deviceNumber (1-based like getDevice(n))
    bool isReady() const { return pExecutor != nullptr; }
    unsigned int getDeviceCount() const { return uiDeviceCount; }
    hp_executor* getDevice() const { return pExecutor; }
    int getDeviceOrdinal() const { return iDeviceOrdinal; }
    std::string getDeviceShortName() const { return sDeviceName; }
    void blockUntilFinished();
  private:
    hp_executor* pExecutor = nullptr;
    int iDeviceOrdinal = 0;
    unsigned int uiDeviceCount = 0;
    std::string sDeviceName;
};

// ---------------------------------------------------------------------------------------------
// Boundaries
// ---------------------------------------------------------------------------------------------
class CBoundary {
  public:
    virtual ~CBoundary() {}
    virtual bool setupFromConfig(const XMLElement* pElement, const std::string& sBoundarySourceDir) = 0;
    virtual void prepareBoundary(hp_scheme* pScheme, CDomainCartesian* pDomain, double dSimulationLength) = 0;   // uploads conf + series
    virtual void importMap(CCSVDataset*) {}
    std::string getName() const { return sName; }
  protected:
    std::string sName;
};

class CBoundaryUniform : public CBoundary {
  public:
    bool setupFromConfig(const XMLElement*, const std::string&) override;
    void prepareBoundary(hp_scheme*, CDomainCartesian*, double) override;
    void importTimeseries(CCSVDataset*);
    std::vector<double> series;            // {t, value} pairs
    unsigned int ucValue = 0;              // 0 rain-intensity, 1 loss-rate
    double dTimeseriesInterval = 0, dTimeseriesLength = 0;
};

class CBoundaryCell : public CBoundary {
  public:
    bool setupFromConfig(const XMLElement*, const std::string&) override;
    void prepareBoundary(hp_scheme*, CDomainCartesian*, double) override;
    void importTimeseries(CCSVDataset*);
    void importMap(CCSVDataset*) override;
    std::vector<double> series;            // {t, depth|fsl, Qx, Qy}
    std::vector<std::pair<unsigned int, unsigned int>> relations;   // cell x, y
    unsigned int ucDepthValue = 1, ucDischargeValue = 1;
    bool bDischargeIsTotal = true;         // kValueTotal == kValuePerCell in the reference (CBoundary.h:54-57)
    double dTimeseriesInterval = 0, dTimeseriesLength = 0;
};

class CBoundaryGridded : public CBoundary {
  public:
    bool setupFromConfig(const XMLElement*, const std::string&) override;
    void prepareBoundary(hp_scheme*, CDomainCartesian*, double) override;
    std::string sMask, sSourceDir;
    double dInterval = 0;
    unsigned int ucValue = 0;              // 0 rain-intensity, 2 mass-flux
};

class CBoundaryMap {
  public:
    bool setupFromConfig(const XMLElement* pXBoundaries, const std::string& sConfigDir);
    void prepareBoundaries(hp_scheme* pScheme, CDomainCartesian* pDomain, double dSimulationLength);
    unsigned int getBoundaryCount() const { return static_cast<unsigned int>(boundaries.size()); }
    CBoundary* getBoundaryByName(const std::string& name);
    std::vector<std::unique_ptr<CBoundary>> boundaries;   // XML order
};

// ---------------------------------------------------------------------------------------------
// Domain
// ---------------------------------------------------------------------------------------------
struct sDataTargetInfo { std::string sValue, sFormat, sTarget; };

class CDomainCartesian {
  public:
    bool configureDomain(const XMLElement* pXDomain, const std::string& sConfigDir);   // structure -> ICs (DEM, depth, others)
    unsigned long getCols() const { return ulCols; }
    unsigned long getRows() const { return ulRows; }
    unsigned long getCellCount() const { return ulCols * ulRows; }
    unsigned long getCellID(unsigned long x, unsigned long y) const { return y * ulCols + x; }
    double getCellResolution() const { return dCellResolution; }
    void handleInputData(unsigned long ulCellID, double dValue, unsigned char ucValue, unsigned char ucRounding);
    static unsigned char getDataValueCode(const std::string& sLower);
    double getVolume() const;
    // derives depth / velocity / fsl / maxdepth / ... rasters: on the device through pScheme when given (one
    // 8-byte value per cell crosses PCIe instead of the whole state), else from the host arrays
    bool writeOutputs(double dTime, CScheme* pScheme = nullptr);
    static double deriveOutput(unsigned char ucValue, const double* state, double bed, double resolution, double nodata);
    // Multi-domain sets (src/Domain/CDomainManager.cpp:56-282, Links/CDomainLink.cpp:73-136): domains stacked north-south
    // with overlapping extents become ONE domain -- a B200's 180 GB holds what the reference spreads over several
    // devices, and the overlap exchange disappears.  Each part is authoritative up to the middle of its overlaps.
    // Returns the merged domain (boundaries moved into it, cell maps shifted) or NULL with doError; `rowOffsets[i]` is
    // the merged row of part i's southern edge.
    static CDomainCartesian* mergeStacked(std::vector<std::unique_ptr<CDomainCartesian>>& parts, std::vector<unsigned long>& rowOffsets);
    // one part's rasters cut out of the merged domain's (north-first) band
    bool writeCroppedOutputs(double dTime, const CDomainCartesian& merged, unsigned long ulRowOffset, CScheme* pScheme,
                             std::map<unsigned char, std::vector<double>>& bandCache) const;
    CBoundaryMap* getBoundaries() { return &boundaryMap; }
    // host cell arrays, reference layout (src/Domain/CDomain.h:28-33); always double on the host side
    std::vector<double> dCellStates;      // cells x {eta, eta_max, qx, qy}
    std::vector<double> dBedElevations, dManningValues;
    std::vector<sDataTargetInfo> outputs;
    std::string sSourceDir, sTargetDir;
    double dRealOffsetX = 0, dRealOffsetY = 0;
  private:
    bool loadInitialConditionSource(unsigned char ucValue, const std::string& type, const std::string& source);
    unsigned long ulCols = 0, ulRows = 0;
    double dCellResolution = 1.0;
    CBoundaryMap boundaryMap;
};

// ---------------------------------------------------------------------------------------------
// Schemes
// ---------------------------------------------------------------------------------------------
class CScheme {
  public:
    static CScheme* createFromConfig(const XMLElement* pXScheme);    // <scheme name="godunov|muscl-hancock|inertial">
    virtual ~CScheme();
    virtual void setupFromConfig(const XMLElement* pXScheme);
    // `devices` (0-based ordinals): more than one runs the domain as row strips, one per GPU, behind this one scheme
    // object -- the multi-device path of the reference (one CScheme per <domain deviceNumber=..>, CDomainLink overlap
    // exchange, src/Domain/CDomainManager.cpp:56-282, Links/CDomainLink.cpp:286-382) re-targeted to the library's NCCL
    // strip engine.  The first device is the executor's.
    bool prepareAll(CExecutorControlCUDA* pExec, CDomainCartesian* pDomain, unsigned char ucFloatPrecision, double dSimulationLength,
                    const std::vector<int>& devices = std::vector<int>());
    unsigned int getStripCount() const { return static_cast<unsigned int>(strips.size()); }
    void prepareSimulation();                          // uploads cells and clock
    void prepareSimulationState();
    void runSimulation(double dTargetTime, double dRealTime);   // sets the target and schedules a batch
    void readKeyStatistics();
    void readDomainAll();                              // device -> CDomain host arrays
    void saveCurrentState() { readDomainAll(); }
    bool deriveRaster(unsigned char ucValue, std::vector<double>& northFirst);   // hp_scheme_derive_raster
    // synchronisation surface of CScheme (src/Schemes/CScheme.h:85-129, CSchemeGodunov.cpp:1474-1616, 1741-1816).  With one
    // merged domain per device there are no link zones, so the rollback limit never binds; the calls keep their meaning.
    bool isSimulationSyncReady(double dExpectedTargetTime) const;       // the target has been reached (to 1e-5 s)
    bool isSimulationFailure(double dExpectedTargetTime) const;         // past the target, or out of rollback budget before it
    // iterations a domain may run between two exchanges of its link zones (src/Domain/CDomainBase.cpp:163-174: the
    // smallest overlap - 1).  The strips of this engine exchange their halo rows in every iteration, so a domain is "not
    // constrained by overlapping" (src/CModel.cpp:543) and keeps the unconstrained value.
    unsigned int getRollbackLimit() const { return uiRollbackLimit; }
    void setRollbackLimit(unsigned int v = 999999999u) { uiRollbackLimit = v; }
    void setSyncMethod(unsigned char m) { ucSyncMethod = m; }
    void rollbackSimulation(double dCurrentTime, double dTargetTime);   // host cell arrays + clock back onto the device
    double proposeSyncPoint(double dCurrentTime) const;
    void forceTimeAdvance() {}                                          // the device clock always advances (no suspended state to leave)
    void setQueueMode(bool bAuto) { bAutomaticQueue = bAuto; }
    bool getQueueMode() const { return bAutomaticQueue; }
    void forceTimestep(double dTimestep);
    void cleanupSimulation();
    bool isReady() const { return pScheme != nullptr; }
    double getCurrentTime() const { return dCurrentTime; }
    double getCurrentTimestep() const { return dCurrentTimestep; }
    unsigned int getIterationsSuccessful() const { return uiBatchSuccessful; }
    unsigned int getIterationsSkipped() const { return uiBatchSkipped; }
    unsigned long long getCellsCalculated() const { return ulCurrentCellsCalculated; }
    unsigned int getBatchSize() const { return uiQueueAdditionSize; }
    double getAverageTimestep() const { return uiBatchSuccessful ? dBatchTimesteps / uiBatchSuccessful : 0.0; }
    // parameters (setters named as in src/Schemes/CScheme.cpp / CSchemeGodunov.cpp)
    void setCourantNumber(double v) { dCourantNumber = v; }
    void setDryThreshold(double v) { dThresholdVerySmall = v; }
    void setTimestepMode(bool dynamic) { bDynamicTimestep = dynamic; }
    void setTimestep(double v) { dTimestep = v; }
    void setFrictionStatus(bool v) { bFrictionEffects = v; }
    void setQueueSize(unsigned int v) { uiQueueAdditionSize = v; }
    void setQuirks(uint32_t q) { uiQuirks = q; }
    void setOptions(uint32_t o) { uiOptions = o; }
    unsigned char getSchemeType() const { return ucSchemeType; }
    hp_scheme* getHandle() const { return pScheme; }
    double dCourantNumber = 0.5, dThresholdVerySmall = 1e-10, dTimestep = 0.001;
    bool bDynamicTimestep = true, bFrictionEffects = true;
  protected:
    explicit CScheme(unsigned char type) : ucSchemeType(type) {}
    unsigned char ucSchemeType;
    unsigned char ucFloatPrecision = model::floatPrecision::kDouble;
    hp_scheme* pScheme = nullptr;                      // the first (on one device: the only) strip
    // one per device: executor (owned here except the first), scheme handle, and the rows it holds / owns
    struct SStrip { hp_executor* pExec = nullptr; bool bOwnsExecutor = false; hp_scheme* pHandle = nullptr;
                    unsigned long ulFirstRow = 0, ulRows = 0, ulOwnFirst = 0, ulOwnRows = 0; };
    std::vector<SStrip> strips;
    template <class F> bool forStrips(F f);            // f(strip, index) on one host thread per strip (NCCL needs them concurrent)
    CDomainCartesian* pDomain = nullptr;
    CExecutorControlCUDA* pExecutor = nullptr;
    unsigned int uiQueueAdditionSize = 256;
    bool bAutomaticQueue = true;                       // CScheme.cpp:46
    unsigned int uiBatchRate = 1;                      // successful iterations of the last batch (CSchemeGodunov.cpp:1834)
    double dBatchStartedTime = 0.0;
    uint32_t uiQuirks = HP_QUIRK_REDUCE_BUFFER_A | HP_QUIRK_BDY_COVERAGE, uiOptions = 0;
    double dCurrentTime = 0, dCurrentTimestep = 0, dBatchTimesteps = 0, dTargetTime = 0;
    unsigned int uiBatchSuccessful = 0, uiBatchSkipped = 0;
    unsigned long long ulCurrentCellsCalculated = 0;
    bool bPeerExchange = false, bPeersAttached = false;     // HIPIMS_STRIP_EXCHANGE=peer: hp_scheme_attach_peers instead of NCCL
    unsigned int uiRollbackLimit = 999999999u;
    unsigned char ucSyncMethod = model::syncMethod::kSyncForecast;
};
class CSchemeGodunov : public CScheme { public: CSchemeGodunov() : CScheme(model::schemeTypes::kGodunov) {} };
class CSchemeMUSCLHancock : public CScheme { public: CSchemeMUSCLHancock() : CScheme(model::schemeTypes::kMUSCLHancock) {} };
class CSchemeInertial : public CScheme { public: CSchemeInertial() : CScheme(model::schemeTypes::kInertialSimplification) {} };

// ---------------------------------------------------------------------------------------------
// Model driver: configuration file -> run to `duration`, writing outputs every `outputFrequency` seconds.  runModel is
// the reference's management loop (src/CModel.cpp:497-527, 552-1139: assess -> rollback -> sync [outputs, new target]
// -> schedule) over ONE local domain, without the polling UI and the MPI hooks; the row strips of a decomposed model
// live inside the domain's scheme and exchange on the device, so "all domains idle / synchronised" is decided from one
// clock.
// ---------------------------------------------------------------------------------------------
class CModel {
  public:
    ~CModel();
    // src/main.cpp:376 + src/Datasets/CXMLDataset.cpp:115-260; bDeviceless parses only (no executor, CPU tests)
    bool loadConfiguration(const std::string& sPath, bool bDeviceless = false);
    bool runModel();
    // the phases of src/CModel.cpp:497-1139, same names and the same state between them
    void runModelPrepare();
    void runModelMain();
    void runModelDomainAssess(bool* bSyncReady, bool* bIdle);
    void runModelDomainExchange();
    void runModelUpdateTarget(double dTimeBase);
    void runModelSync();
    void runModelOutputs();
    void runModelSchedule(double dSeconds, bool* bIdle);
    void runModelRollback();
    void runModelCleanup();
    void writeOutputs();
    unsigned char getSyncMethod() const { return ucSyncMethod; }
    double getCurrentTime() const { return dCurrentTime; }
    double getTargetTime() const { return dTargetTime; }
    double getLastSyncTime() const { return dLastSyncTime; }
    // the target the loop would set after a synchronisation at `dTime` (runModelUpdateTarget on its own: host arithmetic)
    double proposeTargetAfterSyncAt(double dTime) { dLastSyncTime = dTime; dCurrentTime = dTime; runModelUpdateTarget(dTime); return dTargetTime; }
    unsigned int getSyncCount() const { return uiSyncCount; }
    unsigned int getOutputCount() const { return uiOutputCount; }
    // pass wall-clock time to CScheme::runSimulation so that queueMode="auto" sizes batches for about a second of work
    // (src/CModel.cpp:1041-1139); off by default so that iteration counts are reproducible
    void setRealTimeQueue(bool b) { bRealTimeQueue = b; }
    // split the (merged) domain into row strips over these devices (1-based numbers as in <domain deviceNumber=..>);
    // given before loadConfiguration it overrides the configuration's own device numbers (CLI --devices, HIPIMS_DEVICES)
    void setStripDevices(const std::vector<int>& numbers) { stripDeviceNumbers = numbers; }
    unsigned int getStripCount() const { return pScheme ? pScheme->getStripCount() : 0; }
    double getSimulationLength() const { return dSimulationTime; }
    double getOutputFrequency() const { return dOutputFrequency; }
    unsigned char getFloatPrecision() const { return ucFloatPrecision; }
    CDomainCartesian* getDomain() { return pDomain.get(); }
    CScheme* getScheme() { return pScheme.get(); }
    std::string sName, sDescription;
  private:
    bool bRealTimeQueue = false;
    double dSimulationTime = 0, dOutputFrequency = 0;
    // src/CModel.h:105-119
    double dCurrentTime = 0, dLastSyncTime = -1.0, dLastOutputTime = 0, dTargetTime = 0, dEarliestTime = 0, dGlobalTimestep = 0;
    bool bRollbackRequired = false, bAllIdle = true, bWaitOnLinks = false, bSynchronised = true;
    unsigned char ucSyncMethod = model::syncMethod::kSyncForecast;     // <domainSet syncMethod=..>, CDomainManager.cpp:56-76
    unsigned int uiSyncCount = 0, uiOutputCount = 0;
    unsigned char ucFloatPrecision = model::floatPrecision::kDouble;
    std::unique_ptr<CExecutorControlCUDA> pExecutor;
    std::unique_ptr<CDomainCartesian> pDomain;
    std::unique_ptr<CScheme> pScheme;
    // multi-domain configurations: the original domains (for their data targets) and where they sit in pDomain
    std::vector<std::unique_ptr<CDomainCartesian>> parts;
    std::vector<unsigned long> partRowOffsets;
    std::vector<int> stripDeviceNumbers;
  public:
    unsigned int getPartCount() const { return static_cast<unsigned int>(parts.size()); }
    unsigned long getPartRowOffset(unsigned int i) const { return partRowOffsets[i]; }
};
