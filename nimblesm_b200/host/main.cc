// nimblesm_b200/host/main.cc — NimbleSM_b200: the reference's `NimbleSM <input deck>` (src/nimble_main.cc:46-53)
// on the B200 path.
#include "integrator.h"

int
main(int argc, char** argv)
{
  nimble_b200::NimbleApplication app;
  return app.Run(argc, argv);
}
