// oracle/ref/node_tu.cpp -- TEST INFRASTRUCTURE.  Compiles the reference's node.cpp unchanged (by inclusion) and adds the
// one thing a deterministic comparison needs: a way to seed the translation unit's own `generator` (util/randomgen.h:35
// declares it `static`, seeded from std::random_device), which get_dirichlet_noise (util/blazeutil.h:113-124) draws from.
#include "node.cpp"

extern "C" void ref_seed_node_generator(unsigned long long seed) { generator.seed(static_cast<unsigned long>(seed)); }
