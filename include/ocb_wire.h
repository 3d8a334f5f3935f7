/* C ABI of the graph.json wire format for the matching path, exported by libocb_host.so.
 *
 * Replaces, for the members the LinkStage reads and writes, the reference's
 *   bool deserialize(const std::string &json, MeasurementGraph &graph)    src/io/deserialize_MeasurementGraph.cpp:285-288
 *   bool serialize(const MeasurementGraph &graph, std::ostream &out)       src/io/serialize_MeasurementGraph.cpp:592-595
 *   bitset_to_bytes / bitset_from_bytes + Base64encode / Base64decode      serialize :20-27,442-447; deserialize :17-24,
 *                                                                          162-170; src/io/base64.c
 * Plain pointers and sizes only. A graph handle owns everything read from the text; the caller owns every output
 * buffer and sizes it from the *_info calls. All functions return 0 on success and a negative code on failure unless
 * stated otherwise; ocbw_last_error() gives the message of the calling thread's last failure.
 * Layouts: descriptor row = uint64_t[8], the memory image of std::bitset<486> (what ocb_register_descriptors takes);
 * matrices row-major as on the wire; quaternions in Eigen coeffs() order (x, y, z, w).
 */
#ifndef OCB_WIRE_H
#define OCB_WIRE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C"
{
#endif

#define OCBW_DESCRIPTOR_WIRE_BYTES 61
#define OCBW_DESCRIPTOR_BASE64_CHARS 84

    typedef struct ocbw_graph ocbw_graph;

    const char *ocbw_last_error(void);

    /* descriptor row <-> 84 base64 characters (no terminator written / needed) */
    int ocbw_descriptor_encode(const uint64_t row[8], char out[OCBW_DESCRIPTOR_BASE64_CHARS]);
    int ocbw_descriptor_decode(const char *text, size_t n, uint64_t row[8]);
    /* base64 of src/io/base64.c; return the number of bytes written, (size_t)-1 when cap is too small */
    size_t ocbw_base64_encode(const void *bytes, size_t n, char *out, size_t cap);
    size_t ocbw_base64_decode(const char *text, size_t n, void *out, size_t cap);
    /* number <-> text exactly as the reference's rapidjson writer / reader do; buf holds >= 32 characters */
    size_t ocbw_format_double(double value, char *buf);
    int ocbw_parse_double(const char *text, size_t n, double *value);

    /* read: NULL on failure (not a version-1 graph, malformed member) */
    ocbw_graph *ocbw_graph_parse(const char *json, size_t n);
    ocbw_graph *ocbw_graph_create(void); /* empty graph */
    void ocbw_graph_free(ocbw_graph *g);
    size_t ocbw_graph_num_nodes(const ocbw_graph *g);
    size_t ocbw_graph_num_edges(const ocbw_graph *g);

    /* node i in document order. camera[8] = focal_length, principal x/y, radial[3], tangential[2] (the camera
     * model image_to_3d reads); dims[2] = pixels_cols, pixels_rows; pose7 = position[3] + orientation xyzw[4] */
    int ocbw_graph_node_info(const ocbw_graph *g, size_t i, uint64_t *id, size_t *n_features,
                             size_t *num_sparse_features, double camera[8], uint64_t dims[2], double pose7[7]);
    /* features of node i in the device layout: xy [n][2], strength [n], rows [n][8] */
    int ocbw_graph_node_features(const ocbw_graph *g, size_t i, double *xy, float *strength, uint64_t *rows);
    /* append an image node (features in the same layout); path may be NULL. draw_id != 0: the id is drawn like
     * MeasurementGraph::addNode does (graph.hpp:73-84) and returned in *id; otherwise *id is the id to use. */
    int ocbw_graph_add_node(ocbw_graph *g, uint64_t *id, int draw_id, const char *path, const double camera[8],
                            const uint64_t dims[2], const double pose7[7], const double *xy, const float *strength,
                            const uint64_t *rows, size_t n_features, size_t num_sparse_features);

    /* edge i in document order (followed by the edges added since, in the order they were added).
     * relation_type: 0 homography, 1 fundamental_matrix, 2 UNKNOWN; relation[9] row-major;
     * poses[4][8] = score, orientation xyzw, position xyz */
    int ocbw_graph_edge_info(const ocbw_graph *g, size_t i, uint64_t *id, uint64_t *source, uint64_t *dest,
                             size_t *n_matches, size_t *n_inlier_matches, int *relation_type, double relation[9],
                             double poses[32]);
    /* matches: index_1 [n], index_2 [n], distance [n]; inlier_matches: pixels [m][4] (pixel_1 xy, pixel_2 xy),
     * indices [m][3] (feature_index_1, feature_index_2, match_index). Any pointer may be NULL. */
    int ocbw_graph_edge_matches(const ocbw_graph *g, size_t i, uint64_t *index_1, uint64_t *index_2, double *distance,
                                double *inlier_pixels, uint64_t *inlier_indices);
    /* MeasurementGraph::addEdge (include/opencalibration/types/graph.hpp:86-100): stores the relation, draws the
     * edge id like the reference's graph does and registers it with both nodes. *id receives it. */
    int ocbw_graph_add_edge(ocbw_graph *g, uint64_t source, uint64_t dest, const uint64_t *index_1,
                            const uint64_t *index_2, const double *distance, size_t n_matches,
                            const double *inlier_pixels, const uint64_t *inlier_indices, size_t n_inlier_matches,
                            int relation_type, const double relation[9], const double poses[32], uint64_t *id);

    /* write: returns the length of the document; copies it when cap is large enough (no terminator) */
    size_t ocbw_graph_serialize(const ocbw_graph *g, char *out, size_t cap);

    /* LinkStage over the graph's own features (src/pipeline/link_stage.cpp:75-131) on the GPU (needs libocb.so and a
     * device; no CPU fallback): pairs [n_pairs][2] = (node id, neighbour node id). Results become edges in pair
     * order; a pair that already has an edge is replaced in place. run_ransac = 0 stops after the match lists.
     * seconds[4] (nullable) = subsample+upload, match on the GPU, per-pair tail, total. */
    int ocbw_graph_link(ocbw_graph *g, const uint64_t *pairs, size_t n_pairs, int threads, int run_ransac,
                        double seconds[4]);

#ifdef __cplusplus
}
#endif
#endif /* OCB_WIRE_H */
