// =============================================================================
// chrono::dem::ChSystemDem -- source-compatible mirror of the reference class
// (src/chrono_dem/physics/ChSystemDem.h:40-393) on top of the B200 engine's C ABI (include/chrono_b200_dem.h).
// Same method names, argument meaning and order contract (setters -> Initialize -> loop { AdvanceSimulation; getters }).
// Differences: fp64 user units instead of the int32 simulation-unit lattice; the force law
// follows Chrono::Multicore's arithmetic (the Dem user coefficients are mapped); particle
// indices seen through getters/files are always the USER order (the reference re-orders on Initialize,
// src/chrono_dem/gpu/ChDemSMC.cu:185-194, 434-436); errors throw std::runtime_error instead of exit(1).
// =============================================================================
#ifndef CHRONO_B200_CHSYSTEMDEM_H
#define CHRONO_B200_CHSYSTEMDEM_H

#include <climits>
#include <fstream>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include "chrono/core/ChMatrix33.h"
#include "chrono/core/ChQuaternion.h"
#include "chrono/core/ChVector3.h"
#include "chrono/geometry/ChTriangleMeshConnected.h"
#include "chrono_dem/ChApiDem.h"
#include "chrono_dem/ChDemDefines.h"

namespace chrono {
namespace dem {

class ChSystemDem_impl;  // holds the engine handle (dem_b200_system*) and the staged parameters

/// Interface to a Chrono::Dem system.
class CH_DEM_API ChSystemDem {
  public:
    /// Construct system with given sphere radius, density, big domain dimensions and center (ChSystemDem.h:43).
    ChSystemDem(float sphere_rad, float density, const ChVector3f& boxDims, ChVector3f O = ChVector3f(0));
    /// Construct system with a checkpoint file (ChSystemDem.h:46).
    ChSystemDem(const std::string& checkpoint);
    virtual ~ChSystemDem();

    void SetGravitationalAcceleration(const ChVector3f& g);
    void SetParticles(const std::vector<ChVector3f>& points,
                      const std::vector<ChVector3f>& vels = std::vector<ChVector3f>(),
                      const std::vector<ChVector3f>& ang_vels = std::vector<ChVector3f>());
    /// Additive API (hard part 5): per-particle radii for polydisperse packings.  Must match SetParticles.
    void SetParticleRadii(const std::vector<float>& radii);
    void ReadParticleFile(const std::string& infilename);
    void ReadContactHistoryFile(const std::string& infilename);
    void ReadCheckpointFile(const std::string& infilename, bool overwrite = false);

    void SetBDFixed(bool fixed);
    void SetBDCenter(const ChVector3f& O);
    void SetParticleFixed(const std::vector<bool>& fixed);
    void SetParticleOutputMode(CHDEM_OUTPUT_MODE mode);
    void SetParticleOutputFlags(unsigned int flags);
    void SetFixedStepSize(float size_UU);
    float GetFixedStepSize() const;
    void SetDefragmentOnInitialize(bool defragment);
    void EnableMinLength(bool useMinLen);
    void DisableMinLength() { EnableMinLength(false); }
    void SetTimeIntegrator(CHDEM_TIME_INTEGRATOR new_integrator);
    void SetFrictionMode(CHDEM_FRICTION_MODE new_mode);
    void SetRollingMode(CHDEM_ROLLING_MODE new_mode);

    void SetStaticFrictionCoeff_SPH2SPH(float mu);
    void SetStaticFrictionCoeff_SPH2WALL(float mu);
    void SetRollingCoeff_SPH2SPH(float mu);
    void SetRollingCoeff_SPH2WALL(float mu);
    void SetSpinningCoeff_SPH2SPH(float mu);
    void SetSpinningCoeff_SPH2WALL(float mu);
    void SetKn_SPH2SPH(double someValue);
    void SetKn_SPH2WALL(double someValue);
    void SetGn_SPH2SPH(double someValue);
    void SetGn_SPH2WALL(double someValue);
    void SetKt_SPH2SPH(double someValue);
    void SetGt_SPH2SPH(double someValue);
    void SetKt_SPH2WALL(double someValue);
    void SetGt_SPH2WALL(double someValue);
    void SetCohesionRatio(float someValue);
    void SetAdhesionRatio_SPH2WALL(float someValue);
    void UseMaterialBasedModel(bool val);
    void SetYoungModulus_SPH(double someValue);
    void SetYoungModulus_WALL(double someValue);
    void SetPoissonRatio_SPH(double someValue);
    void SetPoissonRatio_WALL(double someValue);
    void SetRestitution_SPH(double someValue);
    void SetRestitution_WALL(double someValue);
    void SetMaxSafeVelocity_SU(float max_vel);
    void SetPsiFactors(unsigned int psi_T, unsigned int psi_L, float psi_R = 1.f);
    void SetPsiT(unsigned int psi_T);
    void SetPsiL(unsigned int psi_L);
    void SetPsiR(float psi_R = 1.f);
    void SetRecordingContactInfo(bool record);
    void SetSimTime(float time);
    void SetVerbosity(CHDEM_VERBOSITY level);

    size_t CreateBCSphere(const ChVector3f& center, float radius, bool outward_normal, bool track_forces, float mass);
    size_t CreateBCConeZ(const ChVector3f& tip, float slope, float hmax, float hmin, bool outward_normal, bool track_forces);
    size_t CreateBCPlane(const ChVector3f& pos, const ChVector3f& normal, bool track_forces);
    size_t CreateCustomizedPlate(const ChVector3f& pos_center, const ChVector3f& normal, float hdim_y);
    size_t CreateBCCylinderZ(const ChVector3f& center, float radius, bool outward_normal, bool track_forces);
    bool DisableBCbyID(size_t BC_id);
    bool EnableBCbyID(size_t BC_id);
    bool SetBCOffsetFunction(size_t BC_id, const GranPositionFunction& offset_function);
    void SetBCSpherePosition(size_t sphere_bc_id, const ChVector3f& pos);
    void SetBCSphereVelocity(size_t sphere_bc_id, const ChVector3f& velo);
    void SetBCPlaneRotation(size_t plane_id, ChVector3d center, ChVector3d omega);
    ChVector3f GetBCSpherePosition(size_t sphere_id) const;
    ChVector3f GetBCSphereVelocity(size_t sphere_id) const;
    void setBDWallsMotionFunction(const GranPositionFunction& pos_fn);

    void SetParticlePosition(int nSphere, const ChVector3d pos);
    void SetParticleDensity(float density);
    void SetParticleRadius(float rad);
    void SetParticleVelocity(int nSphere, const ChVector3d velo);

    float GetSimTime() const;
    size_t GetNumParticles() const;
    double GetMaxParticleZ() const;
    double GetMinParticleZ() const;
    unsigned int GetNumParticleAboveZ(float ZValue) const;
    unsigned int GetNumParticleAboveX(float XValue) const;
    float GetParticleRadius() const;
    ChVector3f GetParticlePosition(int nSphere) const;
    ChVector3f GetParticleVelocity(int nSphere) const;
    ChVector3f GetParticleAngVelocity(int nSphere) const;
    ChVector3f GetParticleLinAcc(int nSphere) const;  ///< declared for source compatibility; throws
    float GetParticlesKineticEnergy() const;
    ChVector3f GetBCPlanePosition(size_t plane_id) const;
    bool IsFixed(int nSphere) const;
    bool GetBCReactionForces(size_t BC_id, ChVector3f& force) const;
    unsigned int GetNumContacts() const;
    unsigned int GetNumSDs() const;

    virtual void Initialize();
    /// Advance simulation by duration in user units, return actual duration elapsed (round(duration/h) steps,
    /// src/chrono_dem/gpu/ChDemSMC.cu:623-624).
    virtual double AdvanceSimulation(float duration);

    void WriteCheckpointFile(const std::string& outfilename);
    void WriteParticleFile(const std::string& outfilename) const;
    void WriteContactHistoryFile(const std::string& outfilename) const;
    /// Per-contact debugging queries of the reference (ChSystemDem.h:320-341): declared for source compatibility; they throw.
    void WriteContactInfoFile(const std::string& outfilename) const;
    ChVector3f getRollingFrictionTorque(unsigned int i, unsigned int j);
    ChVector3f getSlidingFrictionForce(unsigned int i, unsigned int j);
    ChVector3f getNormalForce(unsigned int i, unsigned int j);
    ChVector3f getRollingVrot(unsigned int i, unsigned int j);
    float getRollingCharContactTime(unsigned int i, unsigned int j);
    void getNeighbors(unsigned int ID, std::vector<unsigned int>& neighborList);
    size_t EstimateMemUsage() const;
    float GetRTF() const { return m_RTF; }

    /// Engine handle (dem_b200_system*) for callers that want the C ABI directly (bulk state access, statistics).
    void* GetEngineHandle() const;

  protected:
    ChSystemDem() : m_sys(nullptr), m_RTF(0) {}
    ChSystemDem_impl* m_sys;  ///< underlying system implementation

    void ReadCsvParticles(std::ifstream& ifile, unsigned int totRow = UINT_MAX);
    void ReadHstHistory(std::ifstream& ifile, unsigned int totItem = UINT_MAX);
    virtual bool SetParamsFromIdentifier(const std::string& identifier, std::istringstream& iss1, bool overwrite);
    unsigned int ReadDatParams(std::ifstream& ifile, bool overwrite);
    void WriteCheckpointParams(std::ofstream& cpFile) const;
    void WriteCsvParticles(std::ofstream& ptFile) const;
    void WriteRawParticles(std::ofstream& ptFile) const;
    void WriteHstHistory(std::ofstream& histFile) const;

    float m_RTF;  // real-time factor
};

// -----------------------------------------------------------------------------

/// Interface to a Chrono::Dem mesh system (reference: src/chrono_dem/physics/ChSystemDem.h:398-523).
/// Meshes are rigid bodies carrying one triangle shape per facet; contact follows Chrono::Multicore's
/// triangle_sphere (ChNarrowphasePRIMS.cpp:379-437), which is one-sided.  Chrono::Dem's own test is two-sided
/// (ChDemCollision.cuh:145); to keep that behaviour for existing callers every facet is registered with
/// both windings unless SetMeshTwoSided(false) is called (additive API).
class CH_DEM_API ChSystemDemMesh : public ChSystemDem {
  public:
    ChSystemDemMesh(float sphere_rad, float density, const ChVector3f& boxDims, ChVector3f O = ChVector3f(0));
    ChSystemDemMesh(const std::string& checkpoint);
    ~ChSystemDemMesh();

    unsigned int AddMesh(std::shared_ptr<ChTriangleMeshConnected> mesh, float mass);
    unsigned int AddMesh(const std::string& filename, const ChVector3f& translation, const ChMatrix33<float>& rotscale, float mass);
    std::vector<unsigned int> AddMeshes(const std::vector<std::string>& objfilenames,
                                        const std::vector<ChVector3f>& translations,
                                        const std::vector<ChMatrix33<float>>& rotscales,
                                        const std::vector<float>& masses);
    void EnableMeshCollision(bool val);
    void UseMeshNormals(bool val) { use_mesh_normals = val; }
    /// Additive: false = facets collide only on the side (B-A)x(C-A) points to (Multicore semantics); default true.
    void SetMeshTwoSided(bool val) { m_two_sided = val; }
    void ApplyMeshMotion(unsigned int mesh_id, const ChVector3d& pos, const ChQuaternion<>& rot, const ChVector3d& lin_vel,
                         const ChVector3d& ang_vel);
    unsigned int GetNumMeshes() const;
    std::shared_ptr<ChTriangleMeshConnected> GetMesh(unsigned int mesh_id) const { return m_meshes[mesh_id]; }
    float GetMeshMass(unsigned int mesh_id) const { return m_mesh_masses[mesh_id]; }

    void SetStaticFrictionCoeff_SPH2MESH(float mu);
    void SetRollingCoeff_SPH2MESH(float mu);
    void SetSpinningCoeff_SPH2MESH(float mu);
    void SetKn_SPH2MESH(double someValue);
    void SetGn_SPH2MESH(double someValue);
    void SetKt_SPH2MESH(double someValue);
    void SetGt_SPH2MESH(double someValue);
    void UseMaterialBasedModel(bool val);
    void SetYoungModulus_MESH(double someValue);
    void SetPoissonRatio_MESH(double someValue);
    void SetRestitution_MESH(double someValue);
    void SetAdhesionRatio_SPH2MESH(float someValue);
    void SetMeshVerbosity(CHDEM_MESH_VERBOSITY level) { mesh_verbosity = level; }

    virtual void Initialize() override;
    void InitializeMeshes();
    virtual double AdvanceSimulation(float duration) override;

    void CollectMeshContactForces(std::vector<ChVector3d>& forces, std::vector<ChVector3d>& torques);
    void CollectMeshContactForces(int mesh, ChVector3d& force, ChVector3d& torque);

    void ReadCheckpointFile(const std::string& infilename, bool overwrite = false);
    void WriteCheckpointFile(const std::string& outfilename);
    void WriteMesh(const std::string& outfilename, unsigned int i) const;
    void WriteMeshes(const std::string& outfilename) const;

  private:
    void SetMeshes();
    CHDEM_MESH_VERBOSITY mesh_verbosity = CHDEM_MESH_VERBOSITY::QUIET;
    std::vector<std::shared_ptr<ChTriangleMeshConnected>> m_meshes;
    std::vector<float> m_mesh_masses;
    bool use_mesh_normals = false;
    bool m_two_sided = true;
    virtual bool SetParamsFromIdentifier(const std::string& identifier, std::istringstream& iss1, bool overwrite) override;
    void WriteCheckpointMeshParams(std::ofstream& cpFile) const;
};

}  // namespace dem

// The module was called Chrono::Gpu before the rename; the checkpoint fixture and the CHANGELOG still use that name
// (data/testing/dem/pyramid_checkpoint.dat:1, CHANGELOG.md:3584-3704).
namespace gpu {
using ChSystemGpu = ::chrono::dem::ChSystemDem;
using ChSystemGpuMesh = ::chrono::dem::ChSystemDemMesh;
}
}  // namespace chrono
#endif
