/* A plain C99 host against include/xpbd_fem_b200.h, compiled and run by tests/test_c_host.py (no GPU needed): the boundary is
 * a C ABI - no C++ or torch types - and a scene created host-only (device = -1) can be inspected but never stepped. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "xpbd_fem_b200.h"

int main(void) {
	const uint32_t w = 5, h = 3, d = 2;
	const uint32_t nV = (w + 1) * (h + 1) * (d + 1), nT = 6 * w * h * d;
	float* nodes = (float*)malloc(sizeof(float) * 3 * nV);
	uint32_t* idx = (uint32_t*)malloc(sizeof(uint32_t) * 5 * nT);
	uint32_t* hint = (uint32_t*)malloc(sizeof(uint32_t) * nT);
	uint32_t* order = (uint32_t*)malloc(sizeof(uint32_t) * nT);
	uint32_t* colors = (uint32_t*)malloc(sizeof(uint32_t) * nT);
	uint32_t* chain = (uint32_t*)malloc(sizeof(uint32_t) * nT);
	if (!nodes || !idx || !hint || !order || !colors || !chain) { return 2; }
	if (xf_generate_tet_block(w, h, d, 0.01f, 0.01f, 0.01f, XF_PATTERN_UNIFORM, 0.0f, nodes, idx, hint) != XF_OK) { return 3; }

	xf_create_params p;
	xf_default_create_params(&p);
	if (p.abiVersion != XF_ABI_VERSION) { return 4; }
	p.device = -1; /* host-only */
	p.colorHint = hint;
	p.colorHintCount = nT;
	p.grouping = XF_GROUPING_CHAINS;
	xf_scene* scene = NULL;
	if (xf_create(&p, nodes, 3 * nV, idx, 5 * nT, &scene) != XF_OK) { printf("create failed: %s\n", xf_last_error()); return 5; }
	uint32_t permille = 0;
	if (xf_get_order(scene, order) != XF_OK || xf_get_colors(scene, colors) != XF_OK || xf_get_chain_info(scene, chain, &permille) != XF_OK) { return 6; }
	printf("verts %u elements %u colours %u chained %u\n", xf_vert_count(scene), xf_element_count(scene), xf_color_count(scene), permille);
	/* colour-major order */
	for (uint32_t k = 1; k < nT; k++) {
		if (colors[order[k]] < colors[order[k - 1]]) { return 7; }
	}
	/* there is no CPU compute path: stepping a host-only scene is an error with a message, not a fallback */
	xf_settings st;
	memset(&st, 0, sizeof(st));
	const int rc = xf_substep(scene, &st, NULL, 1.0f / 3000.0f, 1);
	printf("substep rc %d: %s\n", rc, xf_last_error());
	if (rc != XF_ERR_CUDA) { return 8; }
	/* a hint with a conflict is rejected, never silently re-coloured */
	hint[1] = hint[0];
	xf_scene* bad = NULL;
	const int rc2 = xf_create(&p, nodes, 3 * nV, idx, 5 * nT, &bad);
	printf("bad hint rc %d\n", rc2);
	if (rc2 != XF_ERR_COLORING || bad != NULL) { return 9; }
	xf_destroy(scene);
	free(nodes); free(idx); free(hint); free(order); free(colors); free(chain);
	return 0;
}
