// TEST INFRASTRUCTURE ONLY -- the four OpenCV names the reference's graph (de)serializer uses for the node thumbnail
// (serialize_MeasurementGraph.cpp:260-262, deserialize_MeasurementGraph.cpp:78-80). OpenCV's C++ headers are not in
// this image. Thumbnails are outside the matching path: the stand-in encodes every raster as zero PNG bytes and
// decodes to an empty image, so reference-written test documents carry "thumbnail": "".
#pragma once
#include <vector>
typedef unsigned char uchar;
namespace cv
{
struct Mat
{
};
enum
{
    IMREAD_COLOR = 1
};
inline bool imencode(const char *, const Mat &, std::vector<uchar> &out)
{
    out.clear();
    return true;
}
inline void imdecode(const std::vector<uchar> &, int, Mat *) {}
} // namespace cv
