// hipims-b200 -- minimal command line driver: hipims-b200 -c <configuration.xml>
// (the reference's -c option, src/main.cpp:464-499; its UI / logging options are out of scope)
#include <cstdio>
#include <cstring>
#include <string>

#include "hipims_host.h"

int main(int argc, char** argv) {
    std::string config;
    for (int i = 1; i < argc; ++i) {
        if ((!strcmp(argv[i], "-c") || !strcmp(argv[i], "--config-file")) && i + 1 < argc) config = argv[++i];
    }
    if (config.empty()) { fprintf(stderr, "usage: %s -c <configuration.xml>\n", argv[0]); return 2; }
    CModel m;
    if (!m.loadConfiguration(config)) return 1;
    m.setRealTimeQueue(true);                        // queueMode="auto": batches of about a second, like the reference
    printf("%s: %lu x %lu cells, %s, duration %.1f s\n", m.sName.c_str(), m.getDomain()->getCols(), m.getDomain()->getRows(),
           m.getFloatPrecision() == model::floatPrecision::kSingle ? "single" : "double", m.getSimulationLength());
    const double v0 = m.getDomain()->getVolume();
    const bool ok = m.runModel();
    printf("finished at t = %.3f s after %u successful iterations; volume %.3f -> %.3f m3\n", m.getScheme()->getCurrentTime(),
           m.getScheme()->getIterationsSuccessful(), v0, m.getDomain()->getVolume());
    return ok ? 0 : 1;
}
