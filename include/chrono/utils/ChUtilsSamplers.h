// Stand-in for chrono/utils/ChUtilsSamplers.h (reference: src/chrono/utils/ChUtilsSamplers.h), the point samplers the
// Chrono::Dem demos and tests use to create particle clouds:
//   ChGridSampler  regular grid,
//   ChHCPSampler   hexagonally close packed lattice (same lattice as the reference: dx = s, dy = s sqrt(3)/2,
//                  dz = s sqrt(2/3), rows and layers offset; :540-568),
//   ChPDSampler    Poisson-disk sampling (Bridson's algorithm, own implementation; reproducible: fixed seed),
//   ChPDLayerSamplerBox  PD sampling in 2-D layers stacked along z (:454-484).
// Volumes: box, sphere, cylinders along X/Y/Z; a zero half-dimension / half-height gives a 2-D sample.
#ifndef CHRONO_B200_CHUTILSSAMPLERS_H
#define CHRONO_B200_CHUTILSSAMPLERS_H
#include <algorithm>
#include <cmath>
#include <random>
#include <type_traits>
#include <unordered_map>
#include <vector>

#include "chrono/core/ChVector3.h"

namespace chrono {
namespace utils {

enum class SamplingType { REGULAR_GRID, POISSON_DISK, HCP_PACK };

template <typename T>
class ChSampler {
    static_assert(std::is_floating_point<T>::value, "class chrono::utils::ChSampler can only be instantiated with floating point types");

  public:
    typedef std::vector<ChVector3<T>> PointVector;
    virtual ~ChSampler() {}

    PointVector SampleBox(const ChVector3<T>& center, const ChVector3<T>& halfDim) {
        m_center = center; m_size = halfDim;
        return Sample(BOX);
    }
    PointVector SampleSphere(const ChVector3<T>& center, T radius) {
        m_center = center; m_size = ChVector3<T>(radius, radius, radius);
        return Sample(SPHERE);
    }
    PointVector SampleCylinderX(const ChVector3<T>& center, T radius, T halfHeight) {
        m_center = center; m_size = ChVector3<T>(halfHeight, radius, radius);
        return Sample(CYLINDER_X);
    }
    PointVector SampleCylinderY(const ChVector3<T>& center, T radius, T halfHeight) {
        m_center = center; m_size = ChVector3<T>(radius, halfHeight, radius);
        return Sample(CYLINDER_Y);
    }
    PointVector SampleCylinderZ(const ChVector3<T>& center, T radius, T halfHeight) {
        m_center = center; m_size = ChVector3<T>(radius, radius, halfHeight);
        return Sample(CYLINDER_Z);
    }
    virtual T GetSeparation() const { return m_separation; }
    virtual void SetSeparation(T separation) { m_separation = separation; }

  protected:
    enum VolumeType { BOX, SPHERE, CYLINDER_X, CYLINDER_Y, CYLINDER_Z };
    explicit ChSampler(T separation) : m_separation(separation) {}
    virtual PointVector Sample(VolumeType t) = 0;

    /// Is the point inside the current volume (with a small tolerance)?
    bool accept(VolumeType t, const ChVector3<T>& p) const {
        const ChVector3<T> v = p - m_center;
        const T fuzz = (m_size.x() < 1) ? (T)1e-6 * m_size.x() : (T)1e-6;
        auto in = [&](T a, T lim) { return std::abs(a) <= lim + fuzz; };
        switch (t) {
            case BOX: return in(v.x(), m_size.x()) && in(v.y(), m_size.y()) && in(v.z(), m_size.z());
            case SPHERE: return v.Length2() <= m_size.x() * m_size.x();
            case CYLINDER_X: return v.y() * v.y() + v.z() * v.z() <= m_size.y() * m_size.y() && in(v.x(), m_size.x());
            case CYLINDER_Y: return v.z() * v.z() + v.x() * v.x() <= m_size.z() * m_size.z() && in(v.y(), m_size.y());
            case CYLINDER_Z: return v.x() * v.x() + v.y() * v.y() <= m_size.x() * m_size.x() && in(v.z(), m_size.z());
        }
        return false;
    }

    T m_separation;
    ChVector3<T> m_center;
    ChVector3<T> m_size;
};

/// Regular grid.
template <typename T = double>
class ChGridSampler : public ChSampler<T> {
  public:
    typedef typename ChSampler<T>::PointVector PointVector;
    typedef typename ChSampler<T>::VolumeType VolumeType;
    ChGridSampler(T separation) : ChSampler<T>(separation), m_sep3D(separation, separation, separation) {}
    ChGridSampler(const ChVector3<T>& separation) : ChSampler<T>(separation.x()), m_sep3D(separation) {}
    virtual void SetSeparation(T separation) override { m_sep3D = ChVector3<T>(separation, separation, separation); }

  private:
    virtual PointVector Sample(VolumeType t) override {
        PointVector out;
        const ChVector3<T> lo = this->m_center - this->m_size;
        const int nx = (int)(2 * this->m_size.x() / m_sep3D.x()) + 1, ny = (int)(2 * this->m_size.y() / m_sep3D.y()) + 1,
                  nz = (int)(2 * this->m_size.z() / m_sep3D.z()) + 1;
        for (int i = 0; i < nx; i++)
            for (int j = 0; j < ny; j++)
                for (int k = 0; k < nz; k++) {
                    const ChVector3<T> p = lo + ChVector3<T>(i * m_sep3D.x(), j * m_sep3D.y(), k * m_sep3D.z());
                    if (this->accept(t, p))
                        out.push_back(p);
                }
        return out;
    }
    ChVector3<T> m_sep3D;
};

/// Hexagonally close packed lattice.
template <typename T = double>
class ChHCPSampler : public ChSampler<T> {
  public:
    typedef typename ChSampler<T>::PointVector PointVector;
    typedef typename ChSampler<T>::VolumeType VolumeType;
    ChHCPSampler(T separation) : ChSampler<T>(separation) {}

  private:
    virtual PointVector Sample(VolumeType t) override {
        PointVector out;
        const ChVector3<T> lo = this->m_center - this->m_size;
        const T dx = this->m_separation, dy = dx * (T)(std::sqrt(3.0) / 2), dz = dx * (T)std::sqrt(2.0 / 3.0);
        const int nx = (int)(2 * this->m_size.x() / dx) + 1, ny = (int)(2 * this->m_size.y() / dy) + 1,
                  nz = (int)(2 * this->m_size.z() / dz) + 1;
        for (int k = 0; k < nz; k++) {
            const T offy = (k % 2 == 0) ? 0 : dy / 3;          // layers sit in the dimples of the layer below
            for (int j = 0; j < ny; j++) {
                const T offx = ((j + k) % 2 == 0) ? 0 : dx / 2;  // rows alternate by half a separation
                for (int i = 0; i < nx; i++) {
                    const ChVector3<T> p = lo + ChVector3<T>(offx + i * dx, offy + j * dy, k * dz);
                    if (this->accept(t, p))
                        out.push_back(p);
                }
            }
        }
        return out;
    }
};

/// Poisson-disk sampling: no two points closer than the separation, space filled to saturation (Bridson 2007: active
/// list, k candidates per active point in the shell [s, 2s], background grid of cell size s / sqrt(dim)).
template <typename T = double>
class ChPDSampler : public ChSampler<T> {
  public:
    typedef typename ChSampler<T>::PointVector PointVector;
    typedef typename ChSampler<T>::VolumeType VolumeType;
    ChPDSampler(T separation, int pointsPerIteration = 30) : ChSampler<T>(separation), m_k(pointsPerIteration), m_rng(0) {}
    void SetRandomEngineSeed(unsigned int seed) { m_rng.seed(seed); }

  private:
    virtual PointVector Sample(VolumeType t) override {
        const T s = this->m_separation;
        // a vanishing extent makes the sample two-dimensional
        bool flat[3] = {this->m_size.x() < (T)1e-3 * s, this->m_size.y() < (T)1e-3 * s, this->m_size.z() < (T)1e-3 * s};
        if (t == ChSampler<T>::SPHERE) flat[0] = flat[1] = flat[2] = false;
        const int dim = 3 - (int)flat[0] - (int)flat[1] - (int)flat[2];
        PointVector out;
        if (dim == 0) {
            out.push_back(this->m_center);
            return out;
        }
        const T cell = s / (T)std::sqrt((double)dim);
        const ChVector3<T> lo = this->m_center - this->m_size;
        auto key = [&](const ChVector3<T>& p) {
            long long c[3];
            for (unsigned a = 0; a < 3; a++)
                c[a] = flat[a] ? 0 : (long long)std::floor((p[a] - lo[a]) / cell);
            return (c[0] * 73856093LL) ^ (c[1] * 19349663LL) ^ (c[2] * 83492791LL);
        };
        std::unordered_multimap<long long, size_t> grid;
        auto far_enough = [&](const ChVector3<T>& p) {
            long long c[3];
            for (unsigned a = 0; a < 3; a++)
                c[a] = flat[a] ? 0 : (long long)std::floor((p[a] - lo[a]) / cell);
            const int r = 2;
            for (long long i = c[0] - (flat[0] ? 0 : r); i <= c[0] + (flat[0] ? 0 : r); i++)
                for (long long j = c[1] - (flat[1] ? 0 : r); j <= c[1] + (flat[1] ? 0 : r); j++)
                    for (long long k = c[2] - (flat[2] ? 0 : r); k <= c[2] + (flat[2] ? 0 : r); k++) {
                        auto range = grid.equal_range((i * 73856093LL) ^ (j * 19349663LL) ^ (k * 83492791LL));
                        for (auto it = range.first; it != range.second; ++it)
                            if ((out[it->second] - p).Length2() < s * s)
                                return false;
                    }
            return true;
        };
        std::uniform_real_distribution<double> U(0.0, 1.0);
        std::normal_distribution<double> G(0.0, 1.0);
        std::vector<size_t> active;
        auto add = [&](const ChVector3<T>& p) {
            grid.emplace(key(p), out.size());
            active.push_back(out.size());
            out.push_back(p);
        };
        add(this->m_center);
        while (!active.empty()) {
            std::uniform_int_distribution<size_t> pick(0, active.size() - 1);
            const size_t ai = pick(m_rng);
            const ChVector3<T> base = out[active[ai]];
            bool found = false;
            for (int c = 0; c < m_k; c++) {
                // random direction in the free dimensions, radius in [s, 2s]
                double d[3], l2 = 0;
                for (unsigned a = 0; a < 3; a++) {
                    d[a] = flat[a] ? 0.0 : G(m_rng);
                    l2 += d[a] * d[a];
                }
                if (l2 == 0)
                    continue;
                const double rad = (double)s * (1.0 + U(m_rng)) / std::sqrt(l2);
                const ChVector3<T> p = base + ChVector3<T>((T)(d[0] * rad), (T)(d[1] * rad), (T)(d[2] * rad));
                if (!this->accept(t, p) || !far_enough(p))
                    continue;
                add(p);
                found = true;
                break;
            }
            if (!found) {
                active[ai] = active.back();
                active.pop_back();
            }
        }
        return out;
    }

    int m_k;
    std::mt19937 m_rng;
};

/// PD sampling of a box in 2-D layers: layers `padding_factor * diam` apart along z, each a 2-D Poisson-disk sample.
template <typename T>
std::vector<ChVector3<T>> ChPDLayerSamplerBox(ChVector3<T> center, ChVector3<T> hdims, T diam, T padding_factor = 1.02f,
                                              bool verbose = false) {
    (void)verbose;
    const T fill_bottom = center.z() - hdims.z(), fill_top = center.z() + hdims.z();
    ChPDSampler<T> sampler(diam * padding_factor);
    std::vector<ChVector3<T>> out;
    center.z() = fill_bottom;
    hdims.z() = 0;
    unsigned seed = 1;
    while (center.z() < fill_top) {
        sampler.SetRandomEngineSeed(seed++);
        auto pts = sampler.SampleBox(center, hdims);
        out.insert(out.end(), pts.begin(), pts.end());
        center.z() += diam * padding_factor;
    }
    return out;
}

}  // namespace utils
}  // namespace chrono
#endif
