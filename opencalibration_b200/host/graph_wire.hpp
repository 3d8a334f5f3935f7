// Wire format of the matching path: the part of the reference's graph.json checkpoint that the
// LinkStage reads (image nodes with their features and camera model) and writes (edges = camera_relations).
//   reference writer: src/io/serialize_MeasurementGraph.cpp:208-590 (rapidjson PrettyWriter, kFormatSingleLineArray,
//                     kWriteNanAndInfFlag; descriptor = bitset_to_bytes :20-27 -> 61 bytes -> base64 :442-447;
//                     matches as [index_1, index_2, distance] :486-496)
//   reference reader: src/io/deserialize_MeasurementGraph.cpp:31-283 (kParseFullPrecisionFlag | kParseNanAndInfFlag;
//                     bitset_from_bytes :17-24; features :162-177; matches :202-213)
// rapidjson (un-vendored apt dependency of the reference, 1.1.0) and its number formatting (Grisu2 + Prettify,
// rapidjson/internal/dtoa.h) are restated here; base64 follows the reference's src/io/base64.c (APR). The emitter is
// byte-compatible with the reference writer: a document read here and written back is the same byte string
// (the property test/test_serialize_deserialize.cpp:24-64 checks for the reference itself).
// Host code: this is text scanning, not device work; what it produces is the device layout (packed 64-byte
// descriptor rows, SoA locations/strengths) that ocb_register_descriptors / link_pairs consume.
#pragma once
#include "link_batch.hpp"
#include "opencalibration_api.hpp"

#include <array>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <random>
#include <string>
#include <unordered_set>
#include <vector>

namespace ocb_host
{
namespace wire
{
constexpr size_t DESCRIPTOR_WIRE_BYTES = (opencalibration::feature_2d::DESCRIPTOR_BITS + 7) >> 3; // 61
constexpr size_t DESCRIPTOR_BASE64_CHARS = (DESCRIPTOR_WIRE_BYTES + 2) / 3 * 4;                   // 84

// base64 of src/io/base64.c: standard alphabet, '=' padding on encode; decode consumes the longest prefix of
// alphabet characters (stops at '=' or anything else) and drops a dangling single character.
std::string base64_encode(const void *bytes, size_t n);
std::string base64_decode(const char *text, size_t n);

// descriptor row (memory image of std::bitset<486>: 8 little-endian u64, bit j <-> wire byte j>>3 bit j&7)
void descriptor_row_to_base64(const uint64_t row[8], char out[DESCRIPTOR_BASE64_CHARS]);
// false when the text does not decode to exactly 61 bytes (the reference asserts, deserialize :19)
bool descriptor_row_from_base64(const char *text, size_t n, uint64_t row[8]);

// rapidjson Writer::Double (kWriteNanAndInfFlag): shortest-ish round-trip digits by Grisu2, laid out by Prettify
// (maxDecimalPlaces 324), "NaN" / "Infinity" / "-Infinity", "0.0" / "-0.0". Returns the length; buf >= 32 chars.
size_t format_double(double value, char *buf);
// rapidjson ParseNumber with kParseFullPrecisionFlag (correctly rounded) for one complete number token
bool parse_double(const char *text, size_t n, double &value);

// image node: the members of `image` (include/opencalibration/types/image.hpp:18-48) that the checkpoint carries.
// Members outside the matching path (thumbnail, metadata) are kept so that writing the document back loses nothing.
struct CameraInfo
{
    uint64_t width_px = 0, height_px = 0; // defaults of image_metadata (types/image_metadata.hpp:11-21,45-60)
    double focal_length_px = NAN;
    double principal_point_px[2] = {NAN, NAN};
    std::string make, model, serial_no, lens_make, lens_model;
};
struct CaptureInfo
{
    double latitude = NAN, longitude = NAN, altitude = NAN, relative_altitude = NAN, roll = NAN, pitch = NAN,
           yaw = NAN, accuracy_xy = NAN, accuracy_z = NAN;
    std::string datum, timestamp, datestamp;
};
struct GraphNode
{
    size_t id = 0;
    std::string path;
    double position[3] = {NAN, NAN, NAN};              // types/image.hpp:31-32
    double orientation_xyzw[4] = {NAN, NAN, NAN, NAN}; // Eigen coeffs() order, as written (:246-251)
    std::string thumbnail_base64;              // PNG bytes, never decoded here
    int64_t model_id = 0;
    opencalibration::DifferentiableCameraModel<double> model;
    std::string projection = "planar";
    std::vector<size_t> edges; // ids of the incident edges (written sorted, :326-336)
    CameraInfo camera_info;
    CaptureInfo capture_info;
    std::vector<opencalibration::feature_2d> features;
    size_t num_sparse_features = 0;
};
struct GraphEdge
{
    size_t id = 0, source = 0, dest = 0;
    opencalibration::camera_relations relations;
    size_t n_relative_poses = 4; // the reader fills the first rel_pose.Size() entries (:265-281)
};
struct GraphDocument
{
    std::vector<GraphNode> nodes; // document order on read; written sorted by id (:226-233)
    std::vector<GraphEdge> edges;

    const GraphNode *find_node(size_t id) const;
    const GraphEdge *find_edge(size_t source, size_t dest) const;
    // MeasurementGraph::addNode / addEdge (include/opencalibration/types/graph.hpp:73-100): the id is the next draw
    // of the graph's default-seeded std::default_random_engine / uniform_int_distribution<size_t> that is not taken
    // yet (nodes and edges share the generator); addEdge registers the id with both end nodes. A freshly read graph
    // starts with a fresh generator, like a freshly deserialized MeasurementGraph.
    size_t add_node(GraphNode node); // node.id is overwritten by the draw
    size_t add_edge(opencalibration::camera_relations relations, size_t source, size_t dest);
    size_t draw_edge_id(const std::unordered_set<size_t> &taken);

  private:
    std::default_random_engine _generator;
    std::uniform_int_distribution<size_t> _distribution;
};

// deserialize(json, graph) (deserialize_MeasurementGraph.cpp:285-288): false when the text is not a version-1 graph
// object. Unlike the reference (which asserts inside rapidjson) malformed members make it return false with a
// message in *error.
bool read_graph(const char *json, size_t n, GraphDocument &graph, std::string *error = nullptr);
// serialize(graph, out) (serialize_MeasurementGraph.cpp:592-595)
void write_graph(const GraphDocument &graph, std::string &out);

// The LinkStage over a checkpoint: for every pair (source, dest) run what a LinkStage closure runs
// (link_stage.cpp:75-112, batched by link_pairs) on the document's own features and camera models, and store the
// results as edges in the order LinkStage::finalize adds them (:119-131). Pairs that already have an edge are
// replaced in place (same id).
LinkStats link_graph(GraphDocument &graph, const std::vector<LinkPair> &pairs_by_node_id,
                     const LinkOptions &options = LinkOptions());
} // namespace wire
} // namespace ocb_host
