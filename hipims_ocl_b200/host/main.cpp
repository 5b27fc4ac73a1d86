// hipims-b200 -- minimal command line driver: hipims-b200 -c <configuration.xml> [--devices 1,2,3,4]
// (the reference's -c option, src/main.cpp:464-499; its UI / logging options are out of scope).  --devices spreads the
// domain over several GPUs as row strips; without it the <domain deviceNumber=".."> attributes of the configuration decide.
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <string>
#include <vector>

#include "hipims_host.h"

int main(int argc, char** argv) {
    std::string config;
    std::vector<int> devices;
    for (int i = 1; i < argc; ++i) {
        if ((!strcmp(argv[i], "-c") || !strcmp(argv[i], "--config-file")) && i + 1 < argc) config = argv[++i];
        else if (!strcmp(argv[i], "--devices") && i + 1 < argc)
            for (const char* p = argv[++i]; *p;) { devices.push_back(atoi(p)); while (*p && *p != ',') ++p; if (*p) ++p; }
    }
    if (config.empty()) { fprintf(stderr, "usage: %s -c <configuration.xml> [--devices 1,2,...]\n", argv[0]); return 2; }
    CModel m;
    m.setStripDevices(devices);
    if (!m.loadConfiguration(config)) return 1;
    m.setRealTimeQueue(true);                        // queueMode="auto": batches of about a second, like the reference
    printf("%s: %lu x %lu cells, %s, duration %.1f s, %u device(s)\n", m.sName.c_str(), m.getDomain()->getCols(), m.getDomain()->getRows(),
           m.getFloatPrecision() == model::floatPrecision::kSingle ? "single" : "double", m.getSimulationLength(), m.getStripCount());
    const double v0 = m.getDomain()->getVolume();
    const bool ok = m.runModel();
    printf("finished at t = %.3f s after %u successful iterations; volume %.3f -> %.3f m3\n", m.getScheme()->getCurrentTime(),
           m.getScheme()->getIterationsSuccessful(), v0, m.getDomain()->getVolume());
    return ok ? 0 : 1;
}
