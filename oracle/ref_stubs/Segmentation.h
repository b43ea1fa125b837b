/* stand-in for VMD's Segmentation.h / GaussianBlur.h (not in the snapshot): forwards to the
   oracle's restatement ([UNPINNED]).  oracle/_ref build only. */
#ifndef STUB_SEGMENTATION_H
#define STUB_SEGMENTATION_H
template <typename T> class GaussianBlur {
public:
  GaussianBlur(T *img, int w, int h, int d, bool cuda);
  void blur(T sigma);
  T *get_image();
private:
  T *src, *res;
  int nx, ny, nz;
};
#endif
