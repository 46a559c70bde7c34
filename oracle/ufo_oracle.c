/* oracle/ufo_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C, single-threaded CPU restatement of the UFOMap point-cloud
 * integration path (ufo::map::OccupancyMap[Color]::insertPointCloud[Discrete]).
 * It exists to CHECK the CUDA implementation; only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline leg may load it.  The product path
 * (ufomap_b200/) never links, loads or falls back to this file.
 *
 * Parity pin: this restatement is compared node-for-node against the unmodified
 * reference compiled from /root/reference (oracle/_ref/libufo_ref.so, see
 * ref_harness.cpp) by tests/test_oracle_vs_reference.py and against the
 * committed fixtures in tests/golden/ that were generated from that same
 * reference build (tests/golden/make_golden.py).  The reference ships no tests,
 * golden vectors or fixtures of its own (ufomap/tests/CMakeLists.txt is empty).
 *
 * Citations: OMB = ufomap/include/ufo/map/occupancy_map_base.h,
 * OCT = .../octree.h, CODE = .../code.h, OMC.h/.cpp = occupancy_map_color.{h,cpp},
 * V3 = ufomap/include/ufo/math/vector3.h (all under /root/reference).
 *
 * Build with -ffp-contract=off: the reference is built without FMA contraction
 * (CMake default, no -march) and all geometry is sequential IEEE double.
 */
#define _POSIX_C_SOURCE 200809L
#include "ufo_oracle.h"

#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define MAX_LEVELS 22

/* ------------------------------------------------------------------------- */
/* Morton code <-> key  (CODE:183-230: x -> bit 0, y -> bit 1, z -> bit 2)     */
/* ------------------------------------------------------------------------- */
static uint32_t g_spread[256]; /* byte -> bits spaced 3 apart */
static int g_spread_ready = 0;

static void spread_init(void)
{
	if (g_spread_ready) return;
	for (unsigned v = 0; v < 256; ++v) {
		uint32_t s = 0;
		for (unsigned b = 0; b < 8; ++b) s |= ((v >> b) & 1u) << (3 * b);
		g_spread[v] = s;
	}
	g_spread_ready = 1;
}

static uint64_t spread21(uint32_t k)
{
	return (uint64_t)g_spread[k & 0xff] | ((uint64_t)g_spread[(k >> 8) & 0xff] << 24) |
	       ((uint64_t)g_spread[(k >> 16) & 0x1f] << 48);
}

static uint64_t key_to_code(const uint32_t k[3])
{
	return spread21(k[0]) | (spread21(k[1]) << 1) | (spread21(k[2]) << 2);
}

static uint32_t gather21(uint64_t c)
{
	uint32_t k = 0;
	for (unsigned b = 0; b < 21; ++b) k |= (uint32_t)((c >> (3 * b)) & 1u) << b;
	return k;
}

static void code_to_key(uint64_t c, uint32_t k[3])
{
	k[0] = gather21(c);
	k[1] = gather21(c >> 1);
	k[2] = gather21(c >> 2);
}

/* ------------------------------------------------------------------------- */
/* (code, depth) hash set with insertion-ordered iteration                     */
/* Stands in for CodeSet / CodeMap (CODE:378-785): only set semantics matter,  */
/* the reference's bucket iteration order does not affect the value field      */
/* (SURVEY.md Appendix A.7).                                                   */
/* ------------------------------------------------------------------------- */
typedef struct {
	uint64_t* code;
	uint8_t* dp1; /* depth + 1, 0 = empty slot */
	size_t cap, n;
	uint64_t* order_code;
	uint8_t* order_depth;
	size_t order_cap;
} codeset;

static size_t cs_hash(uint64_t c, unsigned d)
{
	uint64_t h = (c + d) * 0x9E3779B97F4A7C15ull;
	return (size_t)(h ^ (h >> 29));
}

static void cs_init(codeset* s, size_t cap_pow2)
{
	memset(s, 0, sizeof(*s));
	s->cap = cap_pow2;
	s->code = (uint64_t*)malloc(s->cap * sizeof(uint64_t));
	s->dp1 = (uint8_t*)calloc(s->cap, 1);
	s->order_cap = 1024;
	s->order_code = (uint64_t*)malloc(s->order_cap * sizeof(uint64_t));
	s->order_depth = (uint8_t*)malloc(s->order_cap);
}

static void cs_free(codeset* s)
{
	free(s->code);
	free(s->dp1);
	free(s->order_code);
	free(s->order_depth);
	memset(s, 0, sizeof(*s));
}

static void cs_place(codeset* s, uint64_t c, unsigned d)
{
	size_t m = s->cap - 1, i = cs_hash(c, d) & m;
	while (s->dp1[i]) i = (i + 1) & m;
	s->code[i] = c;
	s->dp1[i] = (uint8_t)(d + 1);
}

/* returns 1 when newly inserted (like .insert().second / try_emplace().second) */
static int cs_insert(codeset* s, uint64_t c, unsigned d)
{
	size_t m = s->cap - 1, i = cs_hash(c, d) & m;
	while (s->dp1[i]) {
		if (s->code[i] == c && s->dp1[i] == d + 1) return 0;
		i = (i + 1) & m;
	}
	s->code[i] = c;
	s->dp1[i] = (uint8_t)(d + 1);
	if (s->n == s->order_cap) {
		s->order_cap *= 2;
		s->order_code = (uint64_t*)realloc(s->order_code, s->order_cap * sizeof(uint64_t));
		s->order_depth = (uint8_t*)realloc(s->order_depth, s->order_cap);
	}
	s->order_code[s->n] = c;
	s->order_depth[s->n] = (uint8_t)d;
	s->n++;
	if (s->n * 2 > s->cap) {
		free(s->code);
		free(s->dp1);
		s->cap *= 2;
		s->code = (uint64_t*)malloc(s->cap * sizeof(uint64_t));
		s->dp1 = (uint8_t*)calloc(s->cap, 1);
		for (size_t j = 0; j < s->n; ++j) cs_place(s, s->order_code[j], s->order_depth[j]);
	}
	return 1;
}

/* ------------------------------------------------------------------------- */
/* Octree                                                                      */
/* ------------------------------------------------------------------------- */
typedef struct node {
	float occ;       /* log-odds, float32 (LogitType = float, OMB:77)            */
	uint8_t rgb[3];  /* colour payload (color maps only)                         */
	uint8_t cfree;   /* contains_free                                            */
	uint8_t cunk;    /* contains_unknown                                         */
	uint8_t is_leaf; /* no valid children (always 1 for depth-0 voxels)          */
	struct node* ch; /* block of 8 children or NULL                              */
} node;

typedef struct {
	uint64_t code;
	uint32_t depth;
	float occ;
	uint8_t rgb[3];
	uint8_t flags;
} noderec;

typedef struct {
	/* geometry (OCT:922-943) */
	double res, res_factor;
	unsigned levels;
	uint32_t max_value;
	double half[MAX_LEVELS + 2];
	int pruning;
	/* sensor model, double log-odds (OMB:1537-1542) */
	double occ_thr, free_thr, hit, miss, cmin, cmax;
	int color;
	node root;
	size_t n_blocks;
	codeset indices; /* hit / end-voxel dedup, OMB:1552 */
	double min_change[3], max_change[3];
	uint64_t counters[4];
	noderec* scratch;
	size_t scratch_n, scratch_cap;
} omap;

static double to_logit(double p) { return log(p / (1.0 - p)); }

/* OMB:911: toProb takes a float and calls exp on it, i.e. expf. */
static double to_prob(float logit) { return 1.0 / (1.0 + (double)expf(-logit)); }

static int is_free(const omap* m, float v) { return m->free_thr > (double)v; }
static int is_unknown(const omap* m, float v)
{
	return m->free_thr <= (double)v && m->occ_thr >= (double)v;
}

static double node_half(const omap* m, unsigned d) { return m->half[d]; }
static double node_size(const omap* m, unsigned d) { return m->half[d + 1]; }

/* OCT:317-324 */
static uint32_t coord_to_key(const omap* m, double c, unsigned depth)
{
	int kv = (int)floor(m->res_factor * c);
	if (0 == depth) return (uint32_t)kv + m->max_value;
	int snapped = (int)(((unsigned)(kv >> depth)) << depth); /* arithmetic >>, wrap-free << */
	return (uint32_t)(snapped + (1 << (depth - 1))) + m->max_value;
}

static void point_to_key(const omap* m, const double p[3], unsigned depth, uint32_t k[3])
{
	for (int i = 0; i < 3; ++i) k[i] = coord_to_key(m, p[i], depth);
}

/* OCT:372-383 */
static double key_to_coord1(const omap* m, uint32_t key, unsigned depth)
{
	if (m->levels == depth) return 0.0;
	double divider = (double)(1 << depth);
	return (floor(((double)key - (double)m->max_value) / divider) + 0.5) * node_size(m, depth);
}

static void key_to_coord(const omap* m, const uint32_t k[3], unsigned depth, double p[3])
{
	for (int i = 0; i < 3; ++i) p[i] = key_to_coord1(m, k[i], depth);
}

static double v3_sqnorm(const double v[3]) { return (v[0] * v[0]) + (v[1] * v[1]) + (v[2] * v[2]); }
static double v3_norm(const double v[3]) { return sqrt(v3_sqnorm(v)); }

/* ---- BBX clipping, OCT:1240-1332 ---------------------------------------- */
static int in_bbx(const double p[3], double lo, double hi)
{
	return lo <= p[0] && hi >= p[0] && lo <= p[1] && hi >= p[1] && lo <= p[2] && hi >= p[2];
}

static int in_bbx_axis(const double p[3], int axis, double lo, double hi)
{
	int a = (axis + 1) % 3, b = (axis + 2) % 3; /* strict test on the two other axes */
	return p[a] > lo && p[a] < hi && p[b] > lo && p[b] < hi;
}

static int plane_hit(double d1, double d2, const double p1[3], const double p2[3], double hit[3])
{
	if (0 <= (d1 * d2)) return 0;
	double f = -d1 / (d2 - d1);
	for (int i = 0; i < 3; ++i) hit[i] = p1[i] + (p2[i] - p1[i]) * f;
	return 1;
}

static int move_line_inside(const omap* m, double o[3], double e[3])
{
	double hi = node_half(m, m->levels), lo = -hi;
	for (int i = 0; i < 3; ++i) {
		if ((o[i] < lo && e[i] < lo) || (o[i] > hi && e[i] > hi)) return 0;
	}
	if (in_bbx(o, lo, hi) && in_bbx(e, lo, hi)) return 1;

	int hits = 0;
	double hit[2][3];
	for (int i = 0; i < 3 && hits < 2; ++i) {
		if (plane_hit(o[i] - lo, e[i] - lo, o, e, hit[hits]) && in_bbx_axis(hit[hits], i, lo, hi))
			++hits;
	}
	for (int i = 0; i < 3 && hits < 2; ++i) {
		if (plane_hit(o[i] - hi, e[i] - hi, o, e, hit[hits]) && in_bbx_axis(hit[hits], i, lo, hi))
			++hits;
	}
	if (1 == hits) {
		if (in_bbx(o, lo, hi)) memcpy(e, hit[0], sizeof(hit[0]));
		else memcpy(o, hit[0], sizeof(hit[0]));
	} else if (2 == hits) {
		double a0[3], a1[3], b0[3], b1[3];
		for (int i = 0; i < 3; ++i) {
			a0[i] = o[i] - hit[0][i];
			a1[i] = e[i] - hit[1][i];
			b0[i] = o[i] - hit[1][i];
			b1[i] = e[i] - hit[0][i];
		}
		if ((v3_sqnorm(a0) + v3_sqnorm(a1)) <= (v3_sqnorm(b0) + v3_sqnorm(b1))) {
			memcpy(o, hit[0], sizeof(hit[0]));
			memcpy(e, hit[1], sizeof(hit[1]));
		} else {
			memcpy(o, hit[1], sizeof(hit[1]));
			memcpy(e, hit[0], sizeof(hit[0]));
		}
	}
	return 1;
}

/* ---- DDA, OCT:1192-1233 -------------------------------------------------- */
typedef struct {
	uint32_t cur[3], end[3];
	int step[3];
	double t_delta[3], t_max[3];
	int same;
} dda;

static void dda_init(const omap* m, const double a[3], const double b[3], const double dir[3],
                     unsigned depth, dda* s)
{
	point_to_key(m, a, depth, s->cur);
	point_to_key(m, b, depth, s->end);
	s->same = s->cur[0] == s->end[0] && s->cur[1] == s->end[1] && s->cur[2] == s->end[2];
	if (s->same) return;
	double size = node_size(m, depth), hs = node_half(m, depth);
	double border[3];
	key_to_coord(m, s->cur, depth, border);
	for (int i = 0; i < 3; ++i) {
		border[i] = border[i] - a[i];
		if (0 < dir[i]) {
			s->step[i] = (int)(1u << depth);
			border[i] += hs;
			s->t_delta[i] = size / fabs(dir[i]);
			s->t_max[i] = border[i] / dir[i];
		} else if (0 > dir[i]) {
			s->step[i] = -(int)(1u << depth);
			border[i] -= hs;
			s->t_delta[i] = size / fabs(dir[i]);
			s->t_max[i] = border[i] / dir[i];
		} else {
			s->step[i] = 0;
			s->t_delta[i] = DBL_MAX;
			s->t_max[i] = DBL_MAX;
		}
	}
}

/* V3:244-251 tie-break (<=, x before y before z) */
static int min_index(const double t[3])
{
	if (t[0] <= t[1]) return t[0] <= t[2] ? 0 : 2;
	return t[1] <= t[2] ? 1 : 2;
}

static double min3(const double t[3])
{
	double a = t[0] < t[1] ? t[0] : t[1]; /* std::min(std::min(x,y),z) */
	return t[2] < a ? t[2] : a;
}

static void dda_step(dda* s)
{
	int i = min_index(s->t_max);
	s->cur[i] += (uint32_t)s->step[i];
	s->t_max[i] += s->t_delta[i];
}

static int key_eq(const uint32_t a[3], const uint32_t b[3])
{
	return a[0] == b[0] && a[1] == b[1] && a[2] == b[2];
}

/* ---- free space, OMB:1261-1339 -------------------------------------------- */
static void free_space_normal(omap* m, const double from[3], const double to[3], codeset* out,
                              unsigned depth, unsigned early_stopping)
{
	double dir[3], dist;
	for (int i = 0; i < 3; ++i) dir[i] = from[i] - to[i]; /* walked backwards: to -> from */
	dist = v3_norm(dir);
	for (int i = 0; i < 3; ++i) dir[i] /= dist;
	dda s;
	dda_init(m, to, from, dir, depth, &s);
	if (s.same) {
		m->counters[1]++;
		cs_insert(out, key_to_code(s.cur), depth);
		return;
	}
	unsigned in_row = 0;
	do {
		m->counters[1]++;
		if (cs_insert(out, key_to_code(s.cur), depth)) {
			in_row = 0;
		} else {
			++in_row;
			if (0 < early_stopping && in_row >= early_stopping) break;
		}
		dda_step(&s);
	} while (!key_eq(s.cur, s.end) && min3(s.t_max) <= dist);
}

static void free_space_simple(omap* m, const double from[3], const double to[3], codeset* out,
                              unsigned depth, unsigned early_stopping)
{
	double cur[3], dir[3], step[3], dist;
	for (int i = 0; i < 3; ++i) {
		cur[i] = to[i];
		dir[i] = from[i] - to[i];
	}
	dist = v3_norm(dir);
	for (int i = 0; i < 3; ++i) dir[i] /= dist;
	int num_steps = (int)(dist / node_size(m, depth));
	for (int i = 0; i < 3; ++i) step[i] = dir[i] * node_size(m, depth);
	unsigned in_row = 0;
	for (int k = 0; k <= num_steps; ++k) {
		uint32_t key[3];
		point_to_key(m, cur, depth, key);
		m->counters[1]++;
		if (cs_insert(out, key_to_code(key), depth)) {
			in_row = 0;
		} else {
			++in_row;
			if (0 < early_stopping && in_row >= early_stopping) break;
		}
		for (int i = 0; i < 3; ++i) cur[i] += step[i];
	}
}

/* OMB:1229-1259 */
static void free_space(omap* m, const double origin[3], const double* ends, size_t n,
                       codeset* out, unsigned depth, int simple, unsigned early_stopping)
{
	for (size_t r = 0; r < n; ++r) {
		double cur[3] = {origin[0], origin[1], origin[2]};
		double end[3] = {ends[3 * r], ends[3 * r + 1], ends[3 * r + 2]};
		if (!move_line_inside(m, cur, end)) continue;
		m->counters[0]++;
		if (simple) free_space_simple(m, cur, end, out, depth, early_stopping);
		else free_space_normal(m, cur, end, out, depth, early_stopping);
	}
}

/* ---- tree maintenance ------------------------------------------------------ */
static node* alloc_children(omap* m)
{
	node* c = (node*)calloc(8, sizeof(node));
	for (int i = 0; i < 8; ++i) c[i].is_leaf = 1;
	m->n_blocks++;
	return c;
}

/* OCT:1022-1058 */
static void create_children(omap* m, node* n)
{
	if (!n->is_leaf) return;
	if (!n->ch) n->ch = alloc_children(m);
	for (int i = 0; i < 8; ++i) {
		node* keep = n->ch[i].ch;
		n->ch[i] = *n; /* payload, flags and is_leaf(=1) of the parent */
		n->ch[i].ch = keep;
	}
	n->is_leaf = 0;
}

static void free_subtree(omap* m, node* blk, unsigned child_depth)
{
	if (!blk) return;
	if (child_depth > 0) {
		for (int i = 0; i < 8; ++i) free_subtree(m, blk[i].ch, child_depth - 1);
	}
	free(blk);
	m->n_blocks--;
}

/* OCT:1060-1086 */
static void delete_children(omap* m, node* n, unsigned depth)
{
	n->is_leaf = 1;
	if (!n->ch || !m->pruning) return;
	free_subtree(m, n->ch, depth - 1);
	n->ch = NULL;
}

static int payload_equal(const omap* m, const node* a, const node* b)
{
	if (a->occ != b->occ) return 0;
	if (m->color && memcmp(a->rgb, b->rgb, 3) != 0) return 0;
	return 1;
}

/* OCT:1145-1162 */
static int collapsible(const omap* m, const node* n, unsigned depth)
{
	if (1 < depth) {
		for (int i = 0; i < 8; ++i)
			if (!n->ch[i].is_leaf) return 0;
	}
	for (int i = 1; i < 8; ++i)
		if (!payload_equal(m, &n->ch[0], &n->ch[i])) return 0;
	return 1;
}

static int rgb_set(const uint8_t c[3]) { return c[0] || c[1] || c[2]; }

/* OMC.cpp:177-222: root-mean-square of the children whose colour is set */
static void average_child_color(const node* n, uint8_t out[3])
{
	if (n->is_leaf) {
		memcpy(out, n->rgb, 3);
		return;
	}
	double s[3] = {0, 0, 0};
	int cnt = 0;
	for (int i = 0; i < 8; ++i) {
		const uint8_t* c = n->ch[i].rgb;
		if (!rgb_set(c)) continue;
		for (int k = 0; k < 3; ++k) s[k] += (double)c[k] * (double)c[k];
		++cnt;
	}
	if (!cnt) {
		out[0] = out[1] = out[2] = 0;
		return;
	}
	for (int k = 0; k < 3; ++k) out[k] = (uint8_t)sqrt(s[k] / (double)cnt);
}

/* OMB:1179-1224, with the colour override OMC.cpp:115-122 */
static int update_node(omap* m, node* n, unsigned depth)
{
	uint8_t new_rgb[3] = {0, 0, 0};
	if (m->color) average_child_color(n, new_rgb);

	int changed;
	if (n->is_leaf) {
		uint8_t f = (uint8_t)is_free(m, n->occ), u = (uint8_t)is_unknown(m, n->occ);
		changed = (n->cfree != f) || (n->cunk != u);
		n->cfree = f;
		n->cunk = u;
	} else {
		float occ = -FLT_MAX; /* numeric_limits<float>::lowest() */
		uint8_t f = 0, u = 0;
		for (int i = 0; i < 8; ++i) {
			const node* c = &n->ch[i];
			if (c->occ > occ) occ = c->occ;
			if (1 == depth) {
				f = f || is_free(m, c->occ);
				u = u || is_unknown(m, c->occ);
			} else {
				f = f || c->cfree;
				u = u || c->cunk;
			}
		}
		if (collapsible(m, n, depth)) delete_children(m, n, depth);
		changed = (n->occ != occ) || (n->cfree != f) || (n->cunk != u);
		if (changed) {
			n->occ = occ;
			n->cfree = f;
			n->cunk = u;
		}
	}
	if (m->color) {
		changed = changed || memcmp(n->rgb, new_rgb, 3) != 0;
		memcpy(n->rgb, new_rgb, 3);
	}
	return changed;
}

/* OMB:1139-1145: float add, float clamp */
static int update_occupancy(const omap* m, float* cur, float upd)
{
	float old = *cur, v = *cur + upd, lo = (float)m->cmin, hi = (float)m->cmax;
	if (v < lo) v = lo;
	else if (hi < v) v = hi;
	*cur = v;
	return old != v;
}

/* OMB:1085-1120 */
static int update_all_children(omap* m, node* n, unsigned depth, float upd)
{
	int changed = 0;
	for (int i = 0; i < 8; ++i) {
		node* c = &n->ch[i];
		if (1 == depth) {
			if (update_occupancy(m, &c->occ, upd)) changed = 1;
		} else if (c->is_leaf) {
			if (update_occupancy(m, &c->occ, upd)) {
				changed = 1;
				update_node(m, c, depth - 1);
			}
		} else if (update_all_children(m, c, depth - 1, upd)) {
			changed = 1;
		}
	}
	return changed && update_node(m, n, depth);
}

/* OCT:997-1016: descend, materialising children on the way */
static void create_path(omap* m, uint64_t code, unsigned depth, node* path[MAX_LEVELS + 1])
{
	path[m->levels] = &m->root;
	for (unsigned d = m->levels; d > depth; --d) {
		node* n = path[d];
		if (n->is_leaf) create_children(m, n);
		path[d - 1] = &n->ch[(code >> (3 * (d - 1))) & 7u];
	}
}

/* OMB:1126-1133 */
static void update_parents(omap* m, node* path[MAX_LEVELS + 1], unsigned depth)
{
	for (unsigned d = depth > 1 ? depth : 1; d <= m->levels; ++d) {
		if (!update_node(m, path[d], d)) return;
	}
}

/* OMB:1063-1083 */
static void update_value(omap* m, uint64_t code, unsigned depth, float upd)
{
	node* path[MAX_LEVELS + 1];
	create_path(m, code, depth, path);
	if (0 == depth || path[depth]->is_leaf) {
		update_occupancy(m, &path[depth]->occ, upd);
	} else {
		if (!update_all_children(m, path[depth], depth, upd)) return;
		++depth;
	}
	update_parents(m, path, depth);
}

/* OMC.cpp:142-171 */
static void update_leaf_color(node* leaf, const uint8_t upd[3], double prob)
{
	if (0 == memcmp(leaf->rgb, upd, 3)) return;
	if (!rgb_set(leaf->rgb)) {
		memcpy(leaf->rgb, upd, 3);
		return;
	}
	double total = prob + to_prob(leaf->occ);
	prob /= total;
	double inv = 1.0 - prob;
	for (int k = 0; k < 3; ++k) {
		double c = (double)leaf->rgb[k], u = (double)upd[k];
		leaf->rgb[k] = (uint8_t)sqrt(((c * c) * inv) + ((u * u) * prob));
	}
}

/* OMC.h:269-287 */
static void update_value_color(omap* m, uint64_t code, float upd, const uint8_t rgb[3])
{
	node* path[MAX_LEVELS + 1];
	create_path(m, code, 0, path);
	update_leaf_color(path[0], rgb, to_prob(upd));
	update_occupancy(m, &path[0]->occ, upd);
	update_parents(m, path, 0);
}

/* ------------------------------------------------------------------------- */
/* Integration front ends                                                      */
/* ------------------------------------------------------------------------- */
typedef struct {
	uint64_t code;
	uint8_t rgb[3];
} hitrec;

static void bbox_reset(omap* m, double mn[3], double mx[3])
{
	double h = node_half(m, m->levels);
	for (int i = 0; i < 3; ++i) {
		mn[i] = h;  /* min starts at getMax() */
		mx[i] = -h; /* max starts at getMin() */
	}
}

static uint64_t mask_code(const omap* m, uint64_t c)
{
	/* the pointer tree only ever consumes 3*levels code bits (getChildIdx, CODE:245) */
	return m->levels >= 21 ? c : (c & ((1ull << (3 * m->levels)) - 1));
}

/* OMB:1345-1373 / OMC.h:344-373 */
static void integrate(omap* m, const double origin[3], const double* ends, size_t n_ends,
                      const hitrec* hits, size_t n_hits, int hits_have_color, float miss,
                      unsigned depth, int simple, unsigned early_stopping, const double mn[3],
                      const double mx[3])
{
	float hit = (float)m->hit;
	for (size_t i = 0; i < n_hits; ++i) {
		if (hits_have_color) update_value_color(m, mask_code(m, hits[i].code), hit, hits[i].rgb);
		else update_value(m, mask_code(m, hits[i].code), 0, hit);
	}
	codeset free_hits;
	cs_init(&free_hits, 1u << 18);
	free_space(m, origin, ends, n_ends, &free_hits, depth, simple, early_stopping);
	m->counters[2] = free_hits.n;
	m->counters[3] = n_hits;
	for (size_t i = 0; i < free_hits.n; ++i)
		update_value(m, mask_code(m, free_hits.order_code[i]), depth, miss);
	cs_free(&free_hits);
	for (int i = 0; i < 3; ++i) {
		if (mn[i] < m->min_change[i]) m->min_change[i] = mn[i];
		if (mx[i] > m->max_change[i]) m->max_change[i] = mx[i];
	}
}

static void reset_indices(omap* m)
{
	cs_free(&m->indices);
	cs_init(&m->indices, 1u << 18);
}

/* OMB:270-327 (mono) and OMC.h:106-160 (colour branch, which the reference
 * cannot instantiate -- restated from the source text, unpinned) */
static void insert_plain(omap* m, const double so[3], const double* xyz, const uint8_t* rgb,
                         size_t n, double max_range, unsigned depth, int simple,
                         unsigned early_stopping)
{
	hitrec* hits = (hitrec*)malloc((n ? n : 1) * sizeof(hitrec));
	double* ends = (double*)malloc((n ? n : 1) * 3 * sizeof(double));
	size_t nh = 0, ne = 0;
	double mn[3], mx[3];
	bbox_reset(m, mn, mx);
	int use_color = m->color && rgb;
	for (size_t p = 0; p < n; ++p) {
		double end[3] = {xyz[3 * p], xyz[3 * p + 1], xyz[3 * p + 2]};
		double origin[3] = {so[0], so[1], so[2]};
		double dir[3] = {end[0] - origin[0], end[1] - origin[1], end[2] - origin[2]};
		double dist = v3_norm(dir);
		if (!move_line_inside(m, origin, end)) continue;
		if (0 > max_range || dist <= max_range) {
			uint32_t k[3];
			point_to_key(m, end, 0, k);
			uint64_t c = key_to_code(k);
			if (cs_insert(&m->indices, c, 0)) {
				hits[nh].code = c;
				if (use_color) memcpy(hits[nh].rgb, rgb + 3 * p, 3);
				++nh;
			}
		} else {
			for (int i = 0; i < 3; ++i) dir[i] /= dist;
			for (int i = 0; i < 3; ++i) end[i] = origin[i] + (dir[i] * max_range);
		}
		memcpy(ends + 3 * ne, end, sizeof(end));
		++ne;
		for (int i = 0; i < 3; ++i) {
			double lo = end[i] < origin[i] ? end[i] : origin[i];
			double hi = end[i] < origin[i] ? origin[i] : end[i];
			if (lo < mn[i]) mn[i] = lo;
			if (mx[i] < hi) mx[i] = hi;
		}
	}
	float miss = (float)(m->miss / (double)((2.0 * depth) + 1));
	reset_indices(m);
	integrate(m, so, ends, ne, hits, nh, use_color, miss, depth, simple, early_stopping, mn, mx);
	free(hits);
	free(ends);
}

/* OMB:340-417 (mono) and OMC.h:177-267 (colour) */
static void insert_discrete(omap* m, const double so[3], const double* xyz, const uint8_t* rgb,
                            size_t n, double max_range, unsigned depth, int simple,
                            unsigned early_stopping)
{
	hitrec* hits = (hitrec*)malloc((n ? n : 1) * sizeof(hitrec));
	double* ends = (double*)malloc((n ? n : 1) * 3 * sizeof(double));
	size_t nh = 0, ne = 0;
	double mn[3], mx[3];
	bbox_reset(m, mn, mx);
	int use_color = m->color && rgb;
	double sq_max = max_range * max_range;
	double hi = node_half(m, m->levels), lo = -hi;
	for (size_t p = 0; p < n; ++p) {
		double end[3] = {xyz[3 * p], xyz[3 * p + 1], xyz[3 * p + 2]};
		double d[3] = {end[0] - so[0], end[1] - so[1], end[2] - so[2]};
		if (0 > max_range || v3_sqnorm(d) < sq_max) {
			if (in_bbx(end, lo, hi)) {
				uint32_t k[3];
				point_to_key(m, end, 0, k);
				uint64_t c = key_to_code(k);
				if (!cs_insert(&m->indices, c, 0)) continue;
				hits[nh].code = c;
				if (use_color) memcpy(hits[nh].rgb, rgb + 3 * p, 3);
				++nh;
			}
		} else {
			uint32_t k[3];
			double centre[3], dir[3];
			point_to_key(m, end, depth, k);
			key_to_coord(m, k, depth, centre);
			for (int i = 0; i < 3; ++i) dir[i] = centre[i] - so[i];
			if (use_color) {
				/* OMC.h:211-216 compares squared norms */
				double sq = v3_sqnorm(dir);
				if (0 <= max_range && sq > sq_max) {
					double nrm = sqrt(sq);
					for (int i = 0; i < 3; ++i) dir[i] /= nrm;
					for (int i = 0; i < 3; ++i) end[i] = so[i] + (dir[i] * max_range);
				}
			} else {
				/* OMB:364-369 normalises first, compares the norm */
				double nrm = v3_norm(dir);
				for (int i = 0; i < 3; ++i) dir[i] /= nrm;
				if (0 <= max_range && nrm > max_range) {
					for (int i = 0; i < 3; ++i) end[i] = so[i] + (dir[i] * max_range);
				}
			}
		}
		double cur[3] = {so[0], so[1], so[2]};
		if (!move_line_inside(m, cur, end)) continue;
		uint32_t ek[3];
		point_to_key(m, end, depth, ek);
		if (0 < depth && !cs_insert(&m->indices, key_to_code(ek), depth)) continue;
		double ec[3], cc[3];
		uint32_t ck[3];
		key_to_coord(m, ek, depth, ec);
		memcpy(ends + 3 * ne, ec, sizeof(ec));
		++ne;
		point_to_key(m, cur, depth, ck);
		key_to_coord(m, ck, depth, cc);
		double t = node_half(m, depth);
		for (int i = 0; i < 3; ++i) {
			double a = ec[i] - t, b = cc[i] - t;
			double mnv = a < b ? a : b; /* std::min(a, b) */
			if (mnv < mn[i]) mn[i] = mnv;
			a = ec[i] + t;
			b = cc[i] + t;
			double mxv = a < b ? b : a; /* std::max(a, b) */
			if (mx[i] < mxv) mx[i] = mxv;
		}
	}
	float miss = (float)(m->miss / (double)((2.0 * depth) + 1));
	reset_indices(m);
	integrate(m, so, ends, ne, hits, nh, use_color, miss, depth, simple, early_stopping, mn, mx);
	free(hits);
	free(ends);
}

/* ------------------------------------------------------------------------- */
/* C interface                                                                 */
/* ------------------------------------------------------------------------- */
void* ufo_oracle_create(double resolution, unsigned depth_levels, int automatic_pruning,
                        double occupied_thres, double free_thres, double prob_hit,
                        double prob_miss, double clamp_min, double clamp_max, int color)
{
	if (depth_levels < 2 || depth_levels > 21) return NULL; /* OCT:931-935 */
	spread_init();
	omap* m = (omap*)calloc(1, sizeof(omap));
	m->res = resolution;
	m->res_factor = 1.0 / resolution;
	m->levels = depth_levels;
	m->max_value = 1u << (depth_levels - 1);
	m->half[0] = resolution / 2.0;
	m->half[1] = resolution;
	for (unsigned i = 2; i <= depth_levels; ++i) m->half[i] = m->half[i - 1] * 2.0;
	m->pruning = automatic_pruning != 0;
	m->occ_thr = to_logit(occupied_thres);
	m->free_thr = to_logit(free_thres);
	m->hit = to_logit(prob_hit);
	m->miss = to_logit(prob_miss);
	m->cmin = to_logit(clamp_min);
	m->cmax = to_logit(clamp_max);
	m->color = color != 0;
	m->root.is_leaf = 1;
	update_node(m, &m->root, m->levels); /* OMB:871 */
	cs_init(&m->indices, 1u << 18);
	bbox_reset(m, m->min_change, m->max_change);
	return m;
}

void ufo_oracle_destroy(void* h)
{
	omap* m = (omap*)h;
	if (!m) return;
	free_subtree(m, m->root.ch, m->levels - 1);
	cs_free(&m->indices);
	free(m->scratch);
	free(m);
}

double ufo_oracle_insert(void* h, const double* origin, const double* xyz, const uint8_t* rgb,
                         size_t n, double max_range, unsigned depth, int simple,
                         unsigned early_stopping, int discrete, int async_unused)
{
	(void)async_unused;
	omap* m = (omap*)h;
	struct timespec t0, t1;
	clock_gettime(CLOCK_MONOTONIC, &t0);
	memset(m->counters, 0, sizeof(m->counters));
	if (discrete) insert_discrete(m, origin, xyz, rgb, n, max_range, depth, simple, early_stopping);
	else insert_plain(m, origin, xyz, rgb, n, max_range, depth, simple, early_stopping);
	clock_gettime(CLOCK_MONOTONIC, &t1);
	return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

static void scratch_push(omap* m, uint64_t code, unsigned depth, const node* n, uint8_t flags)
{
	if (m->scratch_n == m->scratch_cap) {
		m->scratch_cap = m->scratch_cap ? m->scratch_cap * 2 : 4096;
		m->scratch = (noderec*)realloc(m->scratch, m->scratch_cap * sizeof(noderec));
	}
	noderec* r = &m->scratch[m->scratch_n++];
	r->code = code;
	r->depth = depth;
	r->occ = n->occ;
	memcpy(r->rgb, n->rgb, 3);
	r->flags = flags;
}

static void walk_rec(omap* m, const node* n, unsigned depth, uint64_t code, int leaves)
{
	if ((n->is_leaf != 0) == (leaves != 0))
		scratch_push(m, code, depth, n,
		             (uint8_t)((n->cfree ? 1 : 0) | (n->cunk ? 2 : 0) | (n->is_leaf ? 4 : 0)));
	if (n->is_leaf) return;
	unsigned cd = depth - 1;
	for (unsigned i = 0; i < 8; ++i) {
		uint64_t cc = code + ((uint64_t)i << (3 * cd));
		const node* c = &n->ch[i];
		if (0 == cd) {
			if (leaves)
				scratch_push(m, cc, 0, c,
				             (uint8_t)((is_free(m, c->occ) ? 1 : 0) | (is_unknown(m, c->occ) ? 2 : 0) | 4));
		} else {
			walk_rec(m, c, cd, cc, leaves);
		}
	}
}

size_t ufo_oracle_walk(void* h, int leaves)
{
	omap* m = (omap*)h;
	m->scratch_n = 0;
	walk_rec(m, &m->root, m->levels, 0, leaves);
	return m->scratch_n;
}

void ufo_oracle_walk_fetch(void* h, uint64_t* codes, uint32_t* depths, float* occ, uint8_t* rgb,
                           uint8_t* flags)
{
	omap* m = (omap*)h;
	for (size_t i = 0; i < m->scratch_n; ++i) {
		const noderec* r = &m->scratch[i];
		if (codes) codes[i] = r->code;
		if (depths) depths[i] = r->depth;
		if (occ) occ[i] = r->occ;
		if (rgb) memcpy(rgb + 3 * i, r->rgb, 3);
		if (flags) flags[i] = r->flags;
	}
	free(m->scratch);
	m->scratch = NULL;
	m->scratch_n = m->scratch_cap = 0;
}

/* getNodePath semantics, OCT:957-972 */
int ufo_oracle_node(void* h, uint64_t code, unsigned depth, float* occ, uint8_t* rgb,
                    uint8_t* flags, unsigned* found_depth)
{
	omap* m = (omap*)h;
	const node* n = &m->root;
	unsigned d = m->levels;
	for (; d > depth; --d) {
		if (n->is_leaf) break;
		n = &n->ch[(code >> (3 * (d - 1))) & 7u];
	}
	*occ = n->occ;
	memcpy(rgb, n->rgb, 3);
	if (d > 0) *flags = (uint8_t)((n->cfree ? 1 : 0) | (n->cunk ? 2 : 0) | (n->is_leaf ? 4 : 0));
	else *flags = (uint8_t)((is_free(m, n->occ) ? 1 : 0) | (is_unknown(m, n->occ) ? 2 : 0) | 4);
	*found_depth = d;
	return d == depth;
}

/* OCT:449-496: forward walk, origin voxel included, end voxel excluded */
size_t ufo_oracle_compute_ray(void* h, const double* origin, const double* end_in,
                              double max_range, unsigned depth, uint64_t* codes, size_t cap)
{
	omap* m = (omap*)h;
	double end[3] = {end_in[0], end_in[1], end_in[2]};
	double dir[3] = {end[0] - origin[0], end[1] - origin[1], end[2] - origin[2]};
	double dist = v3_norm(dir);
	for (int i = 0; i < 3; ++i) dir[i] /= dist;
	if (0 <= max_range && dist > max_range) {
		for (int i = 0; i < 3; ++i) end[i] = origin[i] + (dir[i] * max_range);
		dist = max_range;
	}
	dda s;
	dda_init(m, origin, end, dir, depth, &s);
	size_t n = 0;
	if (s.same) return 0;
	while (!key_eq(s.cur, s.end) && min3(s.t_max) <= dist) {
		if (n < cap) codes[n] = key_to_code(s.cur);
		++n;
		dda_step(&s);
	}
	return n;
}

/* castRay, OMB:449-486 -- the reference version does not compile (it tests
 * isOccupied(std::optional) and derives the default range from getMin().distance(getMin())), so
 * this is its INTENT, written down once and shared by the checker and the product:
 *   max_range < 0 -> the map's diagonal; direction normalised; end = origin + direction * max_range;
 *   both clipped into the map (moveLineIntoBBX; outside: no hit); forward walk at `depth` from the
 *   origin's node: an occupied node is returned, an unknown node stops the ray unless
 *   ignore_unknown; the walk ends like computeRay (current == ending or min t_max > max_range) and
 *   the node it stopped in is tested once more.  Node state = the intended getNode semantics (value
 *   of the deepest existing node on the path; inner nodes carry the maximum of their subtree).
 * Returns 1 and the Code (with the depth's centre bits) on a hit. */
int ufo_oracle_cast_ray(void* h, const double* origin_in, const double* dir_in, int ignore_unknown,
                        double max_range, unsigned depth, uint64_t* code)
{
	omap* m = (omap*)h;
	double o[3] = {origin_in[0], origin_in[1], origin_in[2]};
	double d[3] = {dir_in[0], dir_in[1], dir_in[2]};
	if (0 > max_range) {
		double hs = node_half(m, m->levels), diag[3] = {hs + hs, hs + hs, hs + hs};
		max_range = v3_norm(diag);
	}
	double nrm = v3_norm(d);
	for (int i = 0; i < 3; ++i) d[i] /= nrm;
	double e[3];
	for (int i = 0; i < 3; ++i) e[i] = o[i] + (d[i] * max_range);
	if (!move_line_inside(m, o, e)) return 0;
	dda s;
	dda_init(m, o, e, d, depth, &s);
	for (;;) {
		float occ;
		uint8_t rgb[3], fl;
		unsigned fd;
		uint64_t c = key_to_code(s.cur);
		ufo_oracle_node(h, c, depth, &occ, rgb, &fl, &fd);
		int occupied = m->occ_thr < (double)occ;
		int last = s.same || key_eq(s.cur, s.end) || !(min3(s.t_max) <= max_range);
		if (occupied) {
			*code = c;
			return 1;
		}
		if (last) return 0;
		if (!ignore_unknown && is_unknown(m, occ)) return 0;
		dda_step(&s);
	}
}

size_t ufo_oracle_free_set(void* h, const double* origin, const double* ends, size_t n,
                           unsigned depth, int simple, unsigned early_stopping, uint64_t* codes,
                           size_t cap)
{
	omap* m = (omap*)h;
	codeset out;
	cs_init(&out, 1u << 16);
	memset(m->counters, 0, sizeof(m->counters));
	free_space(m, origin, ends, n, &out, depth, simple, early_stopping);
	size_t cnt = out.n;
	for (size_t i = 0; i < cnt && i < cap; ++i) codes[i] = out.order_code[i];
	cs_free(&out);
	return cnt;
}

void ufo_oracle_to_key(void* h, const double* xyz, unsigned depth, uint32_t* key)
{
	point_to_key((const omap*)h, xyz, depth, key);
}

uint64_t ufo_oracle_to_code(void* h, const double* xyz, unsigned depth)
{
	uint32_t k[3];
	point_to_key((const omap*)h, xyz, depth, k);
	return key_to_code(k);
}

uint64_t ufo_oracle_key_to_code(const uint32_t* key, unsigned depth)
{
	(void)depth;
	spread_init();
	return key_to_code(key);
}

void ufo_oracle_code_to_key(uint64_t code, unsigned depth, uint32_t* key)
{
	(void)depth;
	code_to_key(code, key);
}

void ufo_oracle_key_to_coord(void* h, const uint32_t* key, unsigned depth, double* xyz)
{
	key_to_coord((const omap*)h, key, depth, xyz);
}

int ufo_oracle_move_line_inside(void* h, double* a, double* b)
{
	return move_line_inside((const omap*)h, a, b);
}

void ufo_oracle_change_bbox(void* h, double* mn, double* mx)
{
	const omap* m = (const omap*)h;
	memcpy(mn, m->min_change, sizeof(m->min_change));
	memcpy(mx, m->max_change, sizeof(m->max_change));
}

void ufo_oracle_reset_change_bbox(void* h)
{
	omap* m = (omap*)h;
	bbox_reset(m, m->min_change, m->max_change);
}

void ufo_oracle_sensor_model(void* h, double* out6)
{
	const omap* m = (const omap*)h;
	out6[0] = m->occ_thr;
	out6[1] = m->free_thr;
	out6[2] = m->hit;
	out6[3] = m->miss;
	out6[4] = m->cmin;
	out6[5] = m->cmax;
}

size_t ufo_oracle_memory_usage(void* h)
{
	const omap* m = (const omap*)h;
	return sizeof(omap) + m->n_blocks * 8 * sizeof(node);
}

void ufo_oracle_last_counters(void* h, uint64_t* out4)
{
	memcpy(out4, ((const omap*)h)->counters, 4 * sizeof(uint64_t));
}

/* ---- cloud frame (insertPointCloud(..., frame_origin, ...), OMB:313-327) -------------------
 * Pose6::transform  ufomap/include/ufo/math/pose6.h:115-125
 * Quaternion::rotate / operator*  ufomap/include/ufo/math/quaternion.h:253-286
 * Quaternion(roll, pitch, yaw)    ufomap/include/ufo/math/quaternion.h:69-93 */
typedef struct {
	double w, x, y, z;
} quat;

static quat quat_product(quat a, quat b)
{
	quat r;
	r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
	r.x = a.y * b.z - b.y * a.z + a.w * b.x + b.w * a.x;
	r.y = a.z * b.x - b.z * a.x + a.w * b.y + b.w * a.y;
	r.z = a.x * b.y - b.x * a.y + a.w * b.z + b.w * a.z;
	return r;
}

void ufo_oracle_transform(const double* pose7, const double* xyz, size_t n, double* out)
{
	const quat q = {pose7[3], pose7[4], pose7[5], pose7[6]};
	const quat qi = {q.w, -q.x, -q.y, -q.z};
	for (size_t i = 0; i < n; ++i) {
		const quat v = {0.0, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
		const quat r = quat_product(quat_product(q, v), qi);
		out[3 * i + 0] = r.x + pose7[0];
		out[3 * i + 1] = r.y + pose7[1];
		out[3 * i + 2] = r.z + pose7[2];
	}
}

/* std::max(0.0, v) */
static double max0(double v) { return 0.0 < v ? v : 0.0; }

void ufo_oracle_pose_from_rpy(double x, double y, double z, double roll, double pitch, double yaw,
                              double* pose7)
{
	const double sr = sin(roll), sp = sin(pitch), sy = sin(yaw);
	const double cr = cos(roll), cp = cos(pitch), cy = cos(yaw);
	const double m[3][3] = {{cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr},
	                        {sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr},
	                        {-sp, cp * sr, cp * cr}};
	const double w = sqrt(max0(1 + m[0][0] + m[1][1] + m[2][2])) / 2.0;
	const double ax = sqrt(max0(1 + m[0][0] - m[1][1] - m[2][2])) / 2.0;
	const double ay = sqrt(max0(1 - m[0][0] + m[1][1] - m[2][2])) / 2.0;
	const double az = sqrt(max0(1 - m[0][0] - m[1][1] + m[2][2])) / 2.0;
	pose7[0] = x;
	pose7[1] = y;
	pose7[2] = z;
	pose7[3] = w;
	pose7[4] = (m[2][1] - m[1][2]) >= 0 ? fabs(ax) : -fabs(ax);
	pose7[5] = (m[0][2] - m[2][0]) >= 0 ? fabs(ay) : -fabs(ay);
	pose7[6] = (m[1][0] - m[0][1]) >= 0 ? fabs(az) : -fabs(az);
}

/* ---- file image (Octree::write with compress = false, OCT:833-864; writeNodes /
 * writeNodesRecurs OMB:1457-1533; node payloads occupancy_map_node.h:71-75, :106-110, :150-153)
 * header text, then a pre-order walk: one byte per inner node with children (bit i = child i
 * has children), leaf payloads (float occupancy [+ 3 colour bytes]) in child order; the
 * children of depth-1 nodes are written as 8 payloads without a mask byte. */
typedef struct {
	uint8_t* buf;
	size_t cap, n;
} wbuf;

static void wb_put(wbuf* w, const void* p, size_t len)
{
	if (w->buf && w->n + len <= w->cap) memcpy(w->buf + w->n, p, len);
	w->n += len;
}

static void wb_payload(const omap* m, wbuf* w, const node* n)
{
	wb_put(w, &n->occ, 4);
	if (m->color) wb_put(w, n->rgb, 3);
}

static void write_rec(const omap* m, wbuf* w, const node* n, unsigned depth)
{
	const unsigned cd = depth - 1;
	uint8_t children = 0;
	for (unsigned i = 0; i < 8; ++i)
		if (cd > 0 && !n->ch[i].is_leaf) children |= (uint8_t)(1u << i);
	wb_put(w, &children, 1);
	for (unsigned i = 0; i < 8; ++i) {
		const node* c = &n->ch[i];
		if ((children >> i) & 1u) {
			if (1 == cd) {
				for (unsigned j = 0; j < 8; ++j) wb_payload(m, w, &c->ch[j]);
			} else {
				write_rec(m, w, c, cd);
			}
		} else {
			wb_payload(m, w, c);
		}
	}
}

size_t ufo_oracle_write(void* h, uint8_t* buf, size_t cap)
{
	const omap* m = (const omap*)h;
	/* data first (its size goes into the header) */
	wbuf d = {NULL, 0, 0};
	uint8_t children = m->root.is_leaf ? 0 : 0xff;
	for (int pass = 0; pass < 2; ++pass) {
		char head[512];
		int hl = 0;
		if (pass) {
			hl = snprintf(head, sizeof head,
			              "# UFOMap file\n# (feel free to add / change comments, but leave the first line as it "
			              "is!)\n#\nversion 1.0.0\nid %s\nresolution %g\ndepth_levels %u\ncompressed 0\n"
			              "uncompressed_data_size %d\ndata\n",
			              m->color ? "occupancy_map_color" : "occupancy_map", m->res, m->levels, (int)d.n);
		}
		wbuf w = {pass ? buf : NULL, pass ? cap : 0, 0};
		if (pass) wb_put(&w, head, (size_t)hl);
		wb_put(&w, &children, 1);
		if (children) write_rec(m, &w, &m->root, m->levels);
		else wb_payload(m, &w, &m->root);
		if (!pass) d = w;
		else return w.n;
	}
	return 0;
}

/* Test helper: collapse every collapsible node bottom-up (the canonical minimal tree of the map's
 * value field).  The reference's own tree is not always canonical: updateParents stops as soon as
 * an updateNode reports "unchanged" (OMB:1126-1133), so a collapse further up can be missed. */
static void canon_rec(omap* m, node* n, unsigned depth)
{
	if (n->is_leaf) return;
	if (depth > 1)
		for (int i = 0; i < 8; ++i) canon_rec(m, &n->ch[i], depth - 1);
	if (collapsible(m, n, depth)) {
		n->occ = n->ch[0].occ;
		memcpy(n->rgb, n->ch[0].rgb, 3);
		delete_children(m, n, depth);
	}
}

void ufo_oracle_canonicalize(void* h)
{
	omap* m = (omap*)h;
	canon_rec(m, &m->root, m->levels);
}

/* ---- partial / truncated node stream: Octree::writeData(stream, bounding_volume, false,
 * min_depth) (OCT:885-917) with writeNodes / writeNodesRecurs OMB:1457-1533.  Children at
 * min_depth are written as leaves (their aggregate payload); children whose cube misses the
 * bounding box are skipped.  Box test: geometry::intersects(AABB, AABB)
 * (ufomap/src/geometry/collision_checks.cpp:256-264) on getMin()/getMax() of AABB(min, max)
 * (ufomap/include/ufo/geometry/aabb.h:62-69); child centres by getChildCenter (OCT:625-633). */
typedef struct {
	int on;
	double lo[3], hi[3];
} wbox;

static int box_hits(const wbox* bx, const double c[3], double hs)
{
	if (!bx->on) return 1;
	for (int k = 0; k < 3; ++k) {
		const double mn = c[k] - hs, mx = c[k] + hs;
		if (!(mn <= bx->hi[k])) return 0;
		if (!(bx->lo[k] <= mx)) return 0;
	}
	return 1;
}

static void child_center(const double c[3], double hs, unsigned i, double out[3])
{
	out[0] = c[0] + ((i & 1u) ? hs : -hs);
	out[1] = c[1] + ((i & 2u) ? hs : -hs);
	out[2] = c[2] + ((i & 4u) ? hs : -hs);
}

static void write_data_rec(const omap* m, wbuf* w, const node* n, unsigned depth, const double c[3],
                           unsigned min_depth, const wbox* bx)
{
	const unsigned cd = depth - 1;
	const double chs = m->half[cd];
	uint8_t children = 0;
	double cc[8][3];
	int hit[8];
	for (unsigned i = 0; i < 8; ++i) {
		if (cd > min_depth && !n->ch[i].is_leaf) children |= (uint8_t)(1u << i);
		child_center(c, chs, i, cc[i]);
		hit[i] = box_hits(bx, cc[i], chs);
	}
	wb_put(w, &children, 1);
	for (unsigned i = 0; i < 8; ++i) {
		if (!hit[i]) continue;
		const node* ch = &n->ch[i];
		if ((children >> i) & 1u) {
			if (1 == cd) {
				const double ghs = m->half[0];
				for (unsigned j = 0; j < 8; ++j) {
					double gc[3];
					child_center(cc[i], ghs, j, gc);
					if (box_hits(bx, gc, ghs)) wb_payload(m, w, &ch->ch[j]);
				}
			} else {
				write_data_rec(m, w, ch, cd, cc[i], min_depth, bx);
			}
		} else {
			wb_payload(m, w, ch);
		}
	}
}

size_t ufo_oracle_write_data(void* h, const double* box6, unsigned min_depth, uint8_t* buf, size_t cap)
{
	const omap* m = (const omap*)h;
	wbox bx = {0, {0, 0, 0}, {0, 0, 0}};
	if (box6) {
		bx.on = 1;
		for (int k = 0; k < 3; ++k) {
			const double hs = (box6[3 + k] - box6[k]) / 2.0, ct = box6[k] + hs;
			bx.lo[k] = ct - hs;
			bx.hi[k] = ct + hs;
		}
	}
	wbuf w = {buf, cap, 0};
	const double c0[3] = {0.0, 0.0, 0.0};
	if (!box_hits(&bx, c0, m->half[m->levels])) return 0; /* "No node intersects" */
	uint8_t children = (!m->root.is_leaf && m->levels > min_depth) ? 0xff : 0;
	wb_put(&w, &children, 1);
	if (children) write_data_rec(m, &w, &m->root, m->levels, c0, min_depth, &bx);
	else wb_payload(m, &w, &m->root);
	return w.n;
}

/* ---- setValueVolume(AABB, occupancy, min_depth)  OMB:492-518, setValueVolumeRecurs OMB:986-1031,
 * setOccupancy(LogitType&, LogitType const&) OMB:1151-1157 (float clamp of the float logit). */
static int set_occ(const omap* m, float* cur, float v)
{
	const float old = *cur, lo = (float)m->cmin, hi = (float)m->cmax;
	*cur = v < lo ? lo : (hi < v ? hi : v); /* std::clamp<float> */
	return old != *cur;
}

static int set_volume_rec(omap* m, const wbox* bx, float value, node* n, const double c[3], unsigned depth,
                          unsigned min_depth)
{
	const unsigned cd = depth - 1;
	const double chs = m->half[cd];
	create_children(m, n);
	int changed = 0;
	for (unsigned i = 0; i < 8; ++i) {
		double cc[3];
		child_center(c, chs, i, cc);
		if (!box_hits(bx, cc, chs)) continue;
		node* ch = &n->ch[i];
		if (0 == cd) {
			if (set_occ(m, &ch->occ, value)) changed = 1;
		} else if (min_depth < cd) {
			if (set_volume_rec(m, bx, value, ch, cc, cd, min_depth)) changed = 1;
		} else {
			delete_children(m, ch, cd);
			if (set_occ(m, &ch->occ, value)) changed = 1;
			if (update_node(m, ch, cd)) changed = 1;
		}
	}
	return !changed || update_node(m, n, depth);
}

void ufo_oracle_set_value_volume(void* h, const double* box6, double occupancy, unsigned min_depth)
{
	omap* m = (omap*)h;
	if (m->levels < min_depth) return;
	wbox bx = {1, {0, 0, 0}, {0, 0, 0}};
	for (int k = 0; k < 3; ++k) {
		const double hs = (box6[3 + k] - box6[k]) / 2.0, ct = box6[k] + hs;
		bx.lo[k] = ct - hs;
		bx.hi[k] = ct + hs;
	}
	const double c0[3] = {0.0, 0.0, 0.0};
	if (!box_hits(&bx, c0, m->half[m->levels])) return;
	/* the double logit narrows to float at the LogitType parameter of setOccupancy */
	const float value = (float)to_logit(occupancy);
	if (m->levels == min_depth) {
		delete_children(m, &m->root, m->levels);
		set_occ(m, &m->root.occ, value);
		update_node(m, &m->root, m->levels);
		return;
	}
	if (set_volume_rec(m, &bx, value, &m->root, c0, m->levels, min_depth)) update_node(m, &m->root, m->levels);
}

/* ---- readData(stream, bounding volume): readNodes / readNodesRecurs OMB:1379-1455 -- merges a
 * node stream (whole map or the part inside the box it was written with) into the tree. */
typedef struct {
	const uint8_t* p;
	size_t n, at;
} rbuf;

static void rb_get(rbuf* r, void* dst, size_t len)
{
	if (r->at + len <= r->n) memcpy(dst, r->p + r->at, len);
	else memset(dst, 0, len);
	r->at += len;
}

static void rb_payload(const omap* m, rbuf* r, node* n)
{
	rb_get(r, &n->occ, 4);
	if (m->color) rb_get(r, n->rgb, 3);
}

static void read_rec(omap* m, rbuf* r, node* n, unsigned depth, const double c[3], const wbox* bx)
{
	const unsigned cd = depth - 1;
	const double chs = m->half[cd];
	uint8_t children = 0;
	rb_get(r, &children, 1);
	double cc[8][3];
	int hit[8];
	for (unsigned i = 0; i < 8; ++i) {
		child_center(c, chs, i, cc[i]);
		hit[i] = box_hits(bx, cc[i], chs);
	}
	create_children(m, n);
	for (unsigned i = 0; i < 8; ++i) {
		if (!hit[i]) continue;
		node* ch = &n->ch[i];
		if ((children >> i) & 1u) {
			if (1 == cd) {
				const double ghs = m->half[0];
				create_children(m, ch);
				for (unsigned j = 0; j < 8; ++j) {
					double gc[3];
					child_center(cc[i], ghs, j, gc);
					if (box_hits(bx, gc, ghs)) rb_payload(m, r, &ch->ch[j]);
				}
				update_node(m, ch, cd);
			} else {
				read_rec(m, r, ch, cd, cc[i], bx);
			}
		} else {
			delete_children(m, ch, cd);
			rb_payload(m, r, ch);
			update_node(m, ch, cd);
		}
	}
	update_node(m, n, depth);
}

int ufo_oracle_read_data(void* h, const double* box6, const uint8_t* buf, size_t size)
{
	omap* m = (omap*)h;
	wbox bx = {0, {0, 0, 0}, {0, 0, 0}};
	if (box6) {
		bx.on = 1;
		for (int k = 0; k < 3; ++k) {
			const double hs = (box6[3 + k] - box6[k]) / 2.0, ct = box6[k] + hs;
			bx.lo[k] = ct - hs;
			bx.hi[k] = ct + hs;
		}
	}
	const double c0[3] = {0.0, 0.0, 0.0};
	if (!box_hits(&bx, c0, m->half[m->levels])) return 1;
	rbuf r = {buf, size, 0};
	uint8_t children = 0;
	rb_get(&r, &children, 1);
	if (0 == children) {
		delete_children(m, &m->root, m->levels);
		rb_payload(m, &r, &m->root);
		update_node(m, &m->root, m->levels);
		return 1;
	}
	read_rec(m, &r, &m->root, m->levels, c0, &bx);
	return r.at <= r.n;
}
