// TEST INFRASTRUCTURE ONLY. Flat C API over the reference's OWN graph.json reader and writer
//   bool deserialize(const std::string&, MeasurementGraph&)   src/io/deserialize_MeasurementGraph.cpp
//   bool serialize(const MeasurementGraph&, std::ostream&)     src/io/serialize_MeasurementGraph.cpp
//   MeasurementGraph::addNode / addEdge                        include/opencalibration/types/graph.hpp:73-100
// compiled in place from /root/reference into oracle/_ref/liboc_ref_io.so (oracle/Makefile, target ref_io) against the
// storage-only stand-ins of oracle/io_standin and the real rapidjson headers found in this image. Used to generate
// tests/golden/wire_vectors.npz + wire_graph.json (tests/golden/make_wire_vectors.py) and, when present, by
// tests/test_wire.py to pin the product's reader/writer directly. The four raster<->cv conversion functions the two
// translation units reference are defined here as empty thumbnails (see io_standin/opencv2/core.hpp).
#include <opencalibration/io/cv_raster_conversion.hpp>
#include <opencalibration/io/deserialize.hpp>
#include <opencalibration/io/serialize.hpp>

#include <cstring>
#include <sstream>
#include <string>

namespace opencalibration
{
cv::Mat rasterToCv(const GenericRaster &) { return cv::Mat(); }
cv::Mat rasterToCv(const GenericLayer &) { return cv::Mat(); }
GenericRaster cvToRaster(const cv::Mat &) { return GenericRaster(); }
RGBRaster RasterToRGB(const GenericRaster &) { return RGBRaster(); }
} // namespace opencalibration

using namespace opencalibration;

namespace
{
size_t give(const std::string &s, char *out, size_t cap)
{
    if (out && cap >= s.size())
        std::memcpy(out, s.data(), s.size());
    return s.size();
}
} // namespace

extern "C"
{
    // reference reader followed by reference writer; returns the length of the rewritten document, 0 when the
    // reference's deserialize() returns false. *equal_again (nullable) = the rewritten text parses to an equal graph.
    size_t oc_ref_io_roundtrip(const char *json, size_t n, char *out, size_t cap, int *equal_again)
    {
        MeasurementGraph g;
        if (!deserialize(std::string(json, n), g))
            return 0;
        std::ostringstream os;
        serialize(g, os);
        const std::string text = os.str();
        if (equal_again)
        {
            MeasurementGraph h;
            *equal_again = deserialize(text, h) && h == g;
        }
        return give(text, out, cap);
    }

    void *oc_ref_graph_new() { return new MeasurementGraph(); }
    void oc_ref_graph_free(void *g) { delete static_cast<MeasurementGraph *>(g); }

    // addNode with the members the checkpoint carries; strings7 = make, model, serial_no, lens_make, lens_model,
    // datum|timestamp|datestamp are taken from strings[5..7]; capture9 = latitude .. accuracy_z in writer order.
    uint64_t oc_ref_graph_add_node(void *gv, const char *path, const double pose7[7], int64_t model_id,
                                   const double camera8[8], const uint64_t dims2[2], const char *const strings[8],
                                   const double capture9[9], const double *xy, const float *strength,
                                   const uint64_t *rows, size_t n_features, size_t num_sparse)
    {
        MeasurementGraph &g = *static_cast<MeasurementGraph *>(gv);
        image img;
        img.path = path;
        for (int i = 0; i < 3; i++)
            img.position[i] = pose7[i];
        for (int i = 0; i < 4; i++)
            img.orientation.coeffs()[i] = pose7[3 + i];
        img.model = std::make_shared<CameraModel>();
        img.model->id = size_t(model_id);
        img.model->focal_length_pixels = camera8[0];
        img.model->principle_point[0] = camera8[1], img.model->principle_point[1] = camera8[2];
        for (int i = 0; i < 3; i++)
            img.model->radial_distortion[i] = camera8[3 + i];
        img.model->tangential_distortion[0] = camera8[6], img.model->tangential_distortion[1] = camera8[7];
        img.model->pixels_cols = dims2[0], img.model->pixels_rows = dims2[1];
        auto &ci = img.metadata.camera_info;
        ci.width_px = dims2[0], ci.height_px = dims2[1];
        ci.focal_length_px = camera8[0];
        ci.principal_point_px[0] = camera8[1], ci.principal_point_px[1] = camera8[2];
        ci.make = strings[0], ci.model = strings[1], ci.serial_no = strings[2], ci.lens_make = strings[3],
        ci.lens_model = strings[4];
        auto &cap = img.metadata.capture_info;
        cap.latitude = capture9[0], cap.longitude = capture9[1], cap.altitude = capture9[2];
        cap.relativeAltitude = capture9[3], cap.rollDegree = capture9[4], cap.pitchDegree = capture9[5];
        cap.yawDegree = capture9[6], cap.accuracyXY = capture9[7], cap.accuracyZ = capture9[8];
        cap.datum = strings[5], cap.timestamp = strings[6], cap.datestamp = strings[7];
        img.features.resize(n_features);
        for (size_t k = 0; k < n_features; k++)
        {
            img.features[k].location.x() = xy[2 * k], img.features[k].location.y() = xy[2 * k + 1];
            img.features[k].strength = strength[k];
            for (int j = 0; j < feature_2d::DESCRIPTOR_BITS; j++)
                img.features[k].descriptor[j] = (rows[8 * k + (j >> 6)] >> (j & 63)) & 1;
        }
        img.num_sparse_features = num_sparse;
        return g.addNode(std::move(img));
    }

    uint64_t oc_ref_graph_add_edge(void *gv, uint64_t source, uint64_t dest, const uint64_t *i1, const uint64_t *i2,
                                   const double *dist, size_t n_matches, const double *inlier_pixels,
                                   const uint64_t *inlier_idx, size_t n_inliers, int relation_type,
                                   const double relation9[9], const double poses32[32])
    {
        MeasurementGraph &g = *static_cast<MeasurementGraph *>(gv);
        camera_relations rel;
        for (size_t k = 0; k < n_matches; k++)
        {
            feature_match m;
            m.feature_index_1 = i1[k], m.feature_index_2 = i2[k], m.distance = dist[k];
            rel.matches.push_back(m);
        }
        for (size_t k = 0; k < n_inliers; k++)
        {
            feature_match_denormalized m;
            m.pixel_1 = Eigen::Vector2d(inlier_pixels[4 * k], inlier_pixels[4 * k + 1]);
            m.pixel_2 = Eigen::Vector2d(inlier_pixels[4 * k + 2], inlier_pixels[4 * k + 3]);
            m.feature_index_1 = inlier_idx[3 * k], m.feature_index_2 = inlier_idx[3 * k + 1];
            m.match_index = inlier_idx[3 * k + 2];
            rel.inlier_matches.push_back(m);
        }
        rel.relationType = camera_relations::RelationType(relation_type);
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++)
                rel.ransac_relation(r, c) = relation9[3 * r + c];
        for (int p = 0; p < 4; p++)
        {
            rel.relative_poses[p].score = int(poses32[8 * p]);
            for (int k = 0; k < 4; k++)
                rel.relative_poses[p].orientation.coeffs()(k) = poses32[8 * p + 1 + k];
            for (int k = 0; k < 3; k++)
                rel.relative_poses[p].position(k) = poses32[8 * p + 5 + k];
        }
        return g.addEdge(std::move(rel), source, dest);
    }

    size_t oc_ref_graph_serialize(void *gv, char *out, size_t cap)
    {
        std::ostringstream os;
        serialize(*static_cast<MeasurementGraph *>(gv), os);
        return give(os.str(), out, cap);
    }
}
