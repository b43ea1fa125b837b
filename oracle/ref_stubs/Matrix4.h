/* stand-in for VMD's Matrix4 (not in the snapshot): forwards to the oracle's restatement
   (oracle/voltool_oracle.c, [UNPINNED]).  oracle/_ref build only. */
#ifndef STUB_MATRIX4_H
#define STUB_MATRIX4_H
class Matrix4 {
public:
  float mat[16];
  Matrix4();
  Matrix4(const float *m);
  int inverse();
  void multpoint3d(const float in[3], float out[3]) const;
  void rotate_axis(const float axis[3], float angle);
};
#endif
