// Stand-in for chrono/assets/ChVisualSystem.h: the Chrono::Dem demos declare a std::shared_ptr<ChVisualSystem> (and a
// std::shared_ptr<ChBody>) unconditionally and only create one when CHRONO_VSG is defined.  There is no run-time
// visualisation in this repository.
#ifndef CHRONO_B200_CHVISUALSYSTEM_H
#define CHRONO_B200_CHVISUALSYSTEM_H
#include "chrono/core/ChQuaternion.h"
#include "chrono/core/ChVector3.h"
namespace chrono {
/// Proxy body of the run-time visualisation (pose only).
class ChBody {
  public:
    void SetPos(const ChVector3d& p) { m_pos = p; }
    void SetRot(const ChQuaternion<double>& q) { m_rot = q; }
    const ChVector3d& GetPos() const { return m_pos; }
    const ChQuaternion<double>& GetRot() const { return m_rot; }

  private:
    ChVector3d m_pos;
    ChQuaternion<double> m_rot;
};
class ChVisualSystem {
  public:
    virtual ~ChVisualSystem() {}
    virtual bool Run() { return false; }
    virtual void Render() {}
    virtual void BeginScene() {}
    virtual void EndScene() {}
};
}  // namespace chrono
#endif
