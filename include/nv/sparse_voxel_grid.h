// nv/sparse_voxel_grid.h — API-shaped stand-in for the reference's voxel hash
// (libintrinsic3d/include/nv/sparse_voxel_grid.h:69-161, src/sparse_voxel_grid.cpp:166-259).
// Only what the refinement path touches is provided: VoxelSBR, iteration, exists/valid/voxel, voxelSize, truncation.
// Iteration order is insertion order (deterministic); it defines the "voxel_idx" the engine uses.
#pragma once
#include <unordered_map>
#include <utility>
#include <vector>

#include <nv/mat.h>

namespace nv
{
struct VoxelSBR
{
    double sdf = 0.0;
    float weight = 0.0f;
    Vec3b color = Vec3b::Zero();
    double albedo = 0.6;
    double sdf_refined = 0.0;
};

template <class T>
class SparseVoxelGrid
{
public:
    using value_type = std::pair<Vec3i, T>;
    using iterator = typename std::vector<value_type>::iterator;
    using const_iterator = typename std::vector<value_type>::const_iterator;

    static SparseVoxelGrid* create(float voxel_size, float depth_min = 0.1f, float depth_max = 10.0f)
    {
        (void)depth_min; (void)depth_max;
        auto* g = new SparseVoxelGrid();
        g->voxel_size_ = voxel_size;
        g->truncation_ = voxel_size * 5.0f;
        return g;
    }
    iterator begin() { return nodes_.begin(); }
    iterator end() { return nodes_.end(); }
    const_iterator begin() const { return nodes_.begin(); }
    const_iterator end() const { return nodes_.end(); }
    bool empty() const { return nodes_.empty(); }
    size_t numVoxels() const { return nodes_.size(); }
    float voxelSize() const { return voxel_size_; }
    float truncation() const { return truncation_; }

    bool exists(const Vec3i& p) const { return index_.find(p) != index_.end(); }
    bool exists(int x, int y, int z) const { return exists(Vec3i{x, y, z}); }
    bool valid(const Vec3i& p) const { auto it = index_.find(p); return it != index_.end() && nodes_[it->second].second.weight > 0.0f; }
    bool valid(int x, int y, int z) const { return valid(Vec3i{x, y, z}); }
    T& voxel(const Vec3i& p) { return nodes_[index_.find(p)->second].second; }
    const T& voxel(const Vec3i& p) const { return nodes_[index_.find(p)->second].second; }
    T& voxel(int x, int y, int z) { return voxel(Vec3i{x, y, z}); }
    // insertion (the reference fills the grid by TSDF fusion / file load, both out of scope here)
    T& insert(const Vec3i& p, const T& v = T())
    {
        auto it = index_.find(p);
        if (it != index_.end()) { nodes_[it->second].second = v; return nodes_[it->second].second; }
        index_.emplace(p, nodes_.size());
        nodes_.emplace_back(p, v);
        return nodes_.back().second;
    }
    void reserve(size_t n) { nodes_.reserve(n); index_.reserve(n * 2); }
    void clear() { nodes_.clear(); index_.clear(); }
    void setVoxel(const Vec3i& p, const T& v) { insert(p, v); }

private:
    float voxel_size_ = 0.004f, truncation_ = 0.02f;
    std::vector<value_type> nodes_;
    std::unordered_map<Vec3i, size_t> index_;
};
} // namespace nv
