// C ABI of include/ocb_wire.h over host/graph_wire.hpp.
#include "graph_wire.hpp"
#include "ocb_wire.h"

#include <cstring>
#include <exception>
#include <string>

using namespace opencalibration;
namespace w = ocb_host::wire;

struct ocbw_graph
{
    w::GraphDocument doc;
};

namespace
{
thread_local std::string t_err;

template <typename F> int guarded(F &&f)
{
    try
    {
        return f();
    }
    catch (const std::exception &e)
    {
        t_err = e.what();
        return -1;
    }
}
int bad(const char *what)
{
    t_err = what;
    return -2;
}
void fill_features(std::vector<feature_2d> &features, const double *xy, const float *strength, const uint64_t *rows,
                   size_t n)
{
    features.resize(n);
    for (size_t k = 0; k < n; k++)
    {
        features[k].location = Eigen::Vector2d(xy[2 * k], xy[2 * k + 1]);
        features[k].strength = strength[k];
        std::memcpy(static_cast<void *>(&features[k].descriptor), rows + 8 * k, 64);
    }
}
} // namespace

extern "C"
{
    const char *ocbw_last_error(void) { return t_err.c_str(); }

    int ocbw_descriptor_encode(const uint64_t row[8], char out[OCBW_DESCRIPTOR_BASE64_CHARS])
    {
        if (!row || !out)
            return bad("ocbw_descriptor_encode: null argument");
        w::descriptor_row_to_base64(row, out);
        return 0;
    }
    int ocbw_descriptor_decode(const char *text, size_t n, uint64_t row[8])
    {
        if (!text || !row)
            return bad("ocbw_descriptor_decode: null argument");
        if (!w::descriptor_row_from_base64(text, n, row))
            return bad("ocbw_descriptor_decode: the text does not hold 61 base64-coded bytes");
        return 0;
    }
    size_t ocbw_base64_encode(const void *bytes, size_t n, char *out, size_t cap)
    {
        const std::string s = w::base64_encode(bytes, n);
        if (s.size() > cap)
            return size_t(-1);
        std::memcpy(out, s.data(), s.size());
        return s.size();
    }
    size_t ocbw_base64_decode(const char *text, size_t n, void *out, size_t cap)
    {
        const std::string s = w::base64_decode(text, n);
        if (s.size() > cap)
            return size_t(-1);
        std::memcpy(out, s.data(), s.size());
        return s.size();
    }
    size_t ocbw_format_double(double value, char *buf) { return w::format_double(value, buf); }
    int ocbw_parse_double(const char *text, size_t n, double *value)
    {
        return w::parse_double(text, n, *value) ? 0 : bad("ocbw_parse_double: not a JSON number");
    }

    ocbw_graph *ocbw_graph_parse(const char *json, size_t n)
    {
        ocbw_graph *g = nullptr;
        try
        {
            g = new ocbw_graph();
            std::string err;
            if (!json || !w::read_graph(json, n, g->doc, &err))
            {
                t_err = json ? err : "ocbw_graph_parse: null text";
                delete g;
                return nullptr;
            }
            return g;
        }
        catch (const std::exception &e)
        {
            t_err = e.what();
            delete g;
            return nullptr;
        }
    }
    ocbw_graph *ocbw_graph_create(void) { return new ocbw_graph(); }
    void ocbw_graph_free(ocbw_graph *g) { delete g; }
    size_t ocbw_graph_num_nodes(const ocbw_graph *g) { return g ? g->doc.nodes.size() : 0; }
    size_t ocbw_graph_num_edges(const ocbw_graph *g) { return g ? g->doc.edges.size() : 0; }

    int ocbw_graph_node_info(const ocbw_graph *g, size_t i, uint64_t *id, size_t *n_features,
                             size_t *num_sparse_features, double camera[8], uint64_t dims[2], double pose7[7])
    {
        if (!g || i >= g->doc.nodes.size())
            return bad("ocbw_graph_node_info: no such node");
        const w::GraphNode &n = g->doc.nodes[i];
        if (id)
            *id = n.id;
        if (n_features)
            *n_features = n.features.size();
        if (num_sparse_features)
            *num_sparse_features = n.num_sparse_features;
        if (camera)
        {
            camera[0] = n.model.focal_length_pixels;
            camera[1] = n.model.principle_point[0], camera[2] = n.model.principle_point[1];
            for (int k = 0; k < 3; k++)
                camera[3 + k] = n.model.radial_distortion[k];
            camera[6] = n.model.tangential_distortion[0], camera[7] = n.model.tangential_distortion[1];
        }
        if (dims)
            dims[0] = n.model.pixels_cols, dims[1] = n.model.pixels_rows;
        if (pose7)
        {
            std::memcpy(pose7, n.position, sizeof(n.position));
            std::memcpy(pose7 + 3, n.orientation_xyzw, sizeof(n.orientation_xyzw));
        }
        return 0;
    }
    int ocbw_graph_node_features(const ocbw_graph *g, size_t i, double *xy, float *strength, uint64_t *rows)
    {
        if (!g || i >= g->doc.nodes.size())
            return bad("ocbw_graph_node_features: no such node");
        const auto &features = g->doc.nodes[i].features;
        for (size_t k = 0; k < features.size(); k++)
        {
            if (xy)
                xy[2 * k] = features[k].location[0], xy[2 * k + 1] = features[k].location[1];
            if (strength)
                strength[k] = features[k].strength;
            if (rows)
                std::memcpy(rows + 8 * k, static_cast<const void *>(&features[k].descriptor), 64);
        }
        return 0;
    }
    int ocbw_graph_add_node(ocbw_graph *g, uint64_t *id, int draw_id, const char *path, const double camera[8],
                            const uint64_t dims[2], const double pose7[7], const double *xy, const float *strength,
                            const uint64_t *rows, size_t n_features, size_t num_sparse_features)
    {
        if (!g || !id || !camera || !dims || (n_features && (!xy || !strength || !rows)))
            return bad("ocbw_graph_add_node: null argument");
        if (!draw_id && g->doc.find_node(*id))
            return bad("ocbw_graph_add_node: node id already present");
        if (num_sparse_features > n_features)
            return bad("ocbw_graph_add_node: num_sparse_features exceeds n_features");
        return guarded([&] {
            w::GraphNode n;
            n.id = *id;
            n.path = path ? path : "";
            if (pose7)
            {
                std::memcpy(n.position, pose7, sizeof(n.position));
                std::memcpy(n.orientation_xyzw, pose7 + 3, sizeof(n.orientation_xyzw));
            }
            n.model.focal_length_pixels = camera[0];
            n.model.principle_point = Eigen::Vector2d(camera[1], camera[2]);
            n.model.radial_distortion = Eigen::Vector3d(camera[3], camera[4], camera[5]);
            n.model.tangential_distortion = Eigen::Vector2d(camera[6], camera[7]);
            n.model.pixels_cols = dims[0], n.model.pixels_rows = dims[1];
            // the camera_info block carries what the model was initialised from; the rest of the metadata keeps
            // the defaults of image_metadata
            n.camera_info.width_px = dims[0], n.camera_info.height_px = dims[1];
            n.camera_info.focal_length_px = camera[0];
            n.camera_info.principal_point_px[0] = camera[1], n.camera_info.principal_point_px[1] = camera[2];
            fill_features(n.features, xy, strength, rows, n_features);
            n.num_sparse_features = num_sparse_features;
            if (draw_id)
                *id = g->doc.add_node(std::move(n));
            else
                g->doc.nodes.push_back(std::move(n));
            return 0;
        });
    }

    int ocbw_graph_edge_info(const ocbw_graph *g, size_t i, uint64_t *id, uint64_t *source, uint64_t *dest,
                             size_t *n_matches, size_t *n_inlier_matches, int *relation_type, double relation[9],
                             double poses[32])
    {
        if (!g || i >= g->doc.edges.size())
            return bad("ocbw_graph_edge_info: no such edge");
        const w::GraphEdge &e = g->doc.edges[i];
        if (id)
            *id = e.id;
        if (source)
            *source = e.source;
        if (dest)
            *dest = e.dest;
        if (n_matches)
            *n_matches = e.relations.matches.size();
        if (n_inlier_matches)
            *n_inlier_matches = e.relations.inlier_matches.size();
        if (relation_type)
            *relation_type = int(e.relations.relationType);
        if (relation)
            for (int r = 0; r < 3; r++)
                for (int c = 0; c < 3; c++)
                    relation[3 * r + c] = e.relations.ransac_relation(r, c);
        if (poses)
            for (int p = 0; p < 4; p++)
            {
                const decomposed_pose &pose = e.relations.relative_poses[p];
                poses[8 * p] = pose.score;
                for (int k = 0; k < 4; k++)
                    poses[8 * p + 1 + k] = pose.orientation.coeffs()(k);
                for (int k = 0; k < 3; k++)
                    poses[8 * p + 5 + k] = pose.position(k);
            }
        return 0;
    }
    int ocbw_graph_edge_matches(const ocbw_graph *g, size_t i, uint64_t *index_1, uint64_t *index_2, double *distance,
                                double *inlier_pixels, uint64_t *inlier_indices)
    {
        if (!g || i >= g->doc.edges.size())
            return bad("ocbw_graph_edge_matches: no such edge");
        const camera_relations &rel = g->doc.edges[i].relations;
        for (size_t k = 0; k < rel.matches.size(); k++)
        {
            if (index_1)
                index_1[k] = rel.matches[k].feature_index_1;
            if (index_2)
                index_2[k] = rel.matches[k].feature_index_2;
            if (distance)
                distance[k] = rel.matches[k].distance;
        }
        for (size_t k = 0; k < rel.inlier_matches.size(); k++)
        {
            const feature_match_denormalized &m = rel.inlier_matches[k];
            if (inlier_pixels)
            {
                inlier_pixels[4 * k] = m.pixel_1[0], inlier_pixels[4 * k + 1] = m.pixel_1[1];
                inlier_pixels[4 * k + 2] = m.pixel_2[0], inlier_pixels[4 * k + 3] = m.pixel_2[1];
            }
            if (inlier_indices)
            {
                inlier_indices[3 * k] = m.feature_index_1, inlier_indices[3 * k + 1] = m.feature_index_2;
                inlier_indices[3 * k + 2] = m.match_index;
            }
        }
        return 0;
    }
    int ocbw_graph_add_edge(ocbw_graph *g, uint64_t source, uint64_t dest, const uint64_t *index_1,
                            const uint64_t *index_2, const double *distance, size_t n_matches,
                            const double *inlier_pixels, const uint64_t *inlier_indices, size_t n_inlier_matches,
                            int relation_type, const double relation[9], const double poses[32], uint64_t *id)
    {
        if (!g || (n_matches && (!index_1 || !index_2 || !distance)) ||
            (n_inlier_matches && (!inlier_pixels || !inlier_indices)))
            return bad("ocbw_graph_add_edge: null argument");
        if (!g->doc.find_node(source) || !g->doc.find_node(dest))
            return bad("ocbw_graph_add_edge: source or dest is not a node of the graph");
        if (relation_type < 0 || relation_type > 2)
            return bad("ocbw_graph_add_edge: relation_type out of range");
        return guarded([&] {
            camera_relations rel;
            rel.matches.resize(n_matches);
            for (size_t k = 0; k < n_matches; k++)
                rel.matches[k] = feature_match{size_t(index_1[k]), size_t(index_2[k]), distance[k]};
            rel.inlier_matches.resize(n_inlier_matches);
            for (size_t k = 0; k < n_inlier_matches; k++)
            {
                feature_match_denormalized &m = rel.inlier_matches[k];
                m.pixel_1 = Eigen::Vector2d(inlier_pixels[4 * k], inlier_pixels[4 * k + 1]);
                m.pixel_2 = Eigen::Vector2d(inlier_pixels[4 * k + 2], inlier_pixels[4 * k + 3]);
                m.feature_index_1 = inlier_indices[3 * k], m.feature_index_2 = inlier_indices[3 * k + 1];
                m.match_index = inlier_indices[3 * k + 2];
            }
            rel.relationType = camera_relations::RelationType(relation_type);
            if (relation)
                for (int r = 0; r < 3; r++)
                    for (int c = 0; c < 3; c++)
                        rel.ransac_relation(r, c) = relation[3 * r + c];
            if (poses)
                for (int p = 0; p < 4; p++)
                {
                    decomposed_pose &pose = rel.relative_poses[p];
                    pose.score = int(poses[8 * p]);
                    for (int k = 0; k < 4; k++)
                        pose.orientation.coeffs()(k) = poses[8 * p + 1 + k];
                    for (int k = 0; k < 3; k++)
                        pose.position(k) = poses[8 * p + 5 + k];
                }
            const size_t eid = g->doc.add_edge(std::move(rel), source, dest);
            if (id)
                *id = eid;
            return 0;
        });
    }

    size_t ocbw_graph_serialize(const ocbw_graph *g, char *out, size_t cap)
    {
        if (!g)
            return 0;
        std::string text;
        w::write_graph(g->doc, text);
        if (out && cap >= text.size())
            std::memcpy(out, text.data(), text.size());
        return text.size();
    }

    int ocbw_graph_link(ocbw_graph *g, const uint64_t *pairs, size_t n_pairs, int threads, int run_ransac,
                        double seconds[4])
    {
        if (!g || (n_pairs && !pairs))
            return bad("ocbw_graph_link: null argument");
        return guarded([&] {
            std::vector<ocb_host::LinkPair> list(n_pairs);
            for (size_t p = 0; p < n_pairs; p++)
                list[p] = {size_t(pairs[2 * p]), size_t(pairs[2 * p + 1])};
            ocb_host::LinkOptions opt;
            opt.threads = threads;
            opt.run_ransac = run_ransac != 0;
            const ocb_host::LinkStats st = w::link_graph(g->doc, list, opt);
            if (seconds)
            {
                seconds[0] = st.seconds_subsample_upload, seconds[1] = st.seconds_match_gpu;
                seconds[2] = st.seconds_tail, seconds[3] = st.seconds_total;
            }
            return 0;
        });
    }
}
