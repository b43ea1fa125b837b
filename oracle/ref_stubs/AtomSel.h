/* empty stand-in for a VMD header that is not in the reference snapshot (oracle/_ref build only) */
#ifndef STUB_AtomSel_H
#define STUB_AtomSel_H
#endif
