// The C++ assembly shims (include/mx/MxAssembly.hpp) against the C ABI: shapes are host objects and are exercised
// everywhere; the simulation needs a device -- without one its construction must fail loudly (no CPU fallback), with
// one a small problem is assembled end to end and applied.
#include <cmath>
#include <cstdio>

#include "mx/MxAssembly.hpp"

static int fails = 0;
#define EXPECT(cond)                                                      \
  do {                                                                    \
    if (!(cond)) { std::printf("FAILED %s:%d %s\n", __FILE__, __LINE__, #cond); ++fails; } \
  } while (0)

int main() {
  const mx::Vec3 z{{0, 0, 1}}, o{{0, 0, 0}};
  MxShape cyl = MxShape::cylinder(0.4, z, o), caps = MxShape::slab(0.8, z, o);
  MxShape cav = MxShape::intersection({&cyl, &caps});
  EXPECT(cav.func(mx::Vec3{{0, 0, 0}}) > 0);
  EXPECT(cav.func(mx::Vec3{{0.39, 0, 0.39}}) > 0);
  EXPECT(cav.func(mx::Vec3{{0.41, 0, 0}}) < 0);
  EXPECT(cav.func(mx::Vec3{{0, 0, 0.41}}) < 0);
  const mx::Vec3 g = cav.gradFunc(mx::Vec3{{0.3, 0, 0}});
  EXPECT(g[0] < 0 && g[1] == 0 && g[2] == 0);                     // the cylinder is the active constraint there
  MxShape sph = MxShape::sphere(0.2, mx::Vec3{{0.1, 0, 0}});
  sph.translate(mx::Vec3{{0.1, 0, 0}}).scale(mx::Vec3{{2, 1, 1}}, mx::Vec3{{0.2, 0, 0}});
  EXPECT(sph.func(mx::Vec3{{0.55, 0, 0}}) > 0 && sph.func(mx::Vec3{{0.2, 0.25, 0}}) < 0);
  MxShape both = MxShape::unite({&cav, &sph});
  MxShape holed = MxShape::subtract(cav, {&sph});
  EXPECT(both.func(mx::Vec3{{0.55, 0, 0}}) > 0 && holed.func(mx::Vec3{{0.3, 0, 0}}) < 0 && holed.func(mx::Vec3{{-0.3, 0, 0}}) > 0);
  MxShape tilted = MxShape::halfSpace(o, mx::Vec3{{0, 0, 1}});
  tilted.rotate(mx::Vec3{{1, 0, 0}}, 0.5 * std::acos(-1.0));       // the reference's rotation turns by -angle about the axis
  EXPECT(std::fabs(std::fabs(tilted.gradFunc(o)[1]) - 1.0) < 1e-12);
  bool threw = false;
  try { MxShape::intersection({nullptr}); } catch (const std::runtime_error&) { threw = true; }
  EXPECT(threw);

  std::shared_ptr<MxComm> comm;
  try {
    comm = std::make_shared<MxComm>(0);
  } catch (const std::runtime_error& e) {
    std::printf("no device: %s\n", e.what());
    std::printf(fails ? "FAILED\n" : "PASSED (host part)\n");
    return fails ? 1 : 0;
  }
  // a device is present: shape -> operators -> apply
  MxEMSim sim(comm, {{12, 12, 12}}, mx::Vec3{{-0.5, -0.5, -0.5}}, mx::Vec3{{1, 1, 1}});
  sim.setPEC(cav);
  sim.setup();
  auto bmap = sim.getMap("bfield");
  EXPECT(bmap->getNodeNumIndices() == sim.getGlobalIndices("bfield").size());
  MxDeviceCrs ce = sim.getOp("curlE"), cb = sim.getOp("curlB"), dl = sim.getOp("dmL");
  MxDeviceCrs cc = ce.multiply(dl.multiply(cb)), ref = sim.getOp("curlCurl");
  EXPECT(cc.numEntries() == ref.numEntries() && cc.numRows() == ref.numRows());
  auto A = cc.fillComplete<double>(bmap, bmap);
  auto B = ref.fillComplete<double>(bmap, bmap);
  MxMultiVector<double> x(bmap, 2), y1(bmap, 2), y2(bmap, 2);
  x.random();
  A->apply(x, y1);
  B->apply(x, y2);
  y1.update(-1.0, y2, 1.0);
  std::vector<double> nrm(2);
  y1.norm2(nrm);
  EXPECT(nrm[0] == 0.0 && nrm[1] == 0.0);
  std::printf(fails ? "FAILED\n" : "PASSED\n");
  return fails ? 1 : 0;
}
