// Host-side harness over pclean_b200/csrc/lower.hpp (the Plan -> stars lowering is plain C++):
// lets the CPU test-suite check the shape of every block program without a GPU.
#include "../pclean_b200/csrc/lower.hpp"
#include <cstdio>
#include <cstring>

// out[0..8] = {ok, root, n_stars, n_terms, n_roots, rootless, n_root_terms, n_sampled, n_fillins}; err = message on failure
extern "C" int lower_summary(const pclean_model_ir* ir, int cls, int block, int latent, int data_cls,
                             const int* data_obs, int n_data_obs, const int* own_obs, int n_own, int* out, char* err, int err_cap) {
  pcl::Model m; pcl::parse_model(ir, m);
  std::memset(out, 0, 9 * sizeof(int));
  try {
    pcl::Lowerer L(m, cls);
    L.intern = [&](const std::u32string& s) { m.strings.push_back(s); return (int)m.strings.size() - 1; };
    std::vector<char> dobs(m.classes[data_cls].nv, 0);
    for (int i = 0; i < n_data_obs; ++i) dobs[data_obs[i]] = 1;
    std::vector<char> own(m.classes[cls].nv, 0);
    for (int i = 0; i < n_own; ++i) own[own_obs[i]] = 1;
    if (latent) { L.latent = true; L.data_cls = data_cls; L.data_obs = &dobs; L.ir = ir; }
    pcl::BlockProgram p = L.lower_block(block, latent ? own : dobs);
    int fills = 0;
    for (auto& s : p.stars) fills += (int)s.fillins.size();
    out[0] = 1; out[1] = p.root; out[2] = (int)p.stars.size(); out[3] = (int)p.terms.size(); out[4] = (int)p.roots.size();
    out[5] = p.rootless ? 1 : 0; out[6] = (int)p.root_terms.size(); out[7] = (int)p.root_sampled.size(); out[8] = fills;
    return 0;
  } catch (const std::exception& e) {
    std::snprintf(err, err_cap, "%s", e.what());
    return -1;
  }
}
