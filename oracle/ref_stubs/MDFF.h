/* declaration of the one function MDFF.C defines (MDFF.C:102) (oracle/_ref build only) */
#ifndef STUB_MDFF_H
#define STUB_MDFF_H
class VolumetricData;
int cc_threaded(VolumetricData *qsVol, const VolumetricData *targetVol, double *cc, double threshold);
#endif
