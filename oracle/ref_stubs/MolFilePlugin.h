/* empty stand-in for a VMD header that is not in the reference snapshot (oracle/_ref build only) */
#ifndef STUB_MolFilePlugin_H
#define STUB_MolFilePlugin_H
#endif
