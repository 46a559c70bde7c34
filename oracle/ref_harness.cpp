// oracle/ref_harness.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Thin C-ABI shim around the UNMODIFIED reference (UFOMap, headers included from
// /root/reference at compile time; nothing is copied into this repo).  It is
// built into oracle/_ref/libufo_ref.so by oracle/Makefile and is used
//   * by tests/ to pin the CPU restatement (oracle/ufo_oracle.c) and the CUDA
//     path against the real reference implementation, and to generate the
//     committed fixtures in tests/golden/ (tests/golden/make_golden.py);
//   * by bench.py's cpu_baseline / --impl reference leg (kind = "reference").
// Nothing in the product path (ufomap_b200/) may link or load this file.
//
// State is read through a subclass that walks the pointer octree directly
// (root -> children), NOT through getOccupancy(code): Octree::getNode has an
// off-by-one (octree.h:974-985) and returns the depth-1 parent.

#include <ufo/map/occupancy_map.h>
#include <ufo/map/occupancy_map_color.h>

#include <chrono>
#include <cstdint>
#include <cstring>
#include <sstream>
#include <string>
#include <vector>

// ---- LZ4 (the image has liblz4.so.1 without development files): the four entry points the
// reference calls (octree.h:1436-1445, :1473) forward to the system library, loaded on first use
#include <dlfcn.h>
namespace
{
void* lz4_sym(const char* name)
{
	static void* lib = dlopen("liblz4.so.1", RTLD_NOW | RTLD_LOCAL);
	return lib ? dlsym(lib, name) : nullptr;
}
}  // namespace
extern "C" {
int LZ4_compressBound(int n)
{
	auto f = reinterpret_cast<int (*)(int)>(lz4_sym("LZ4_compressBound"));
	return f ? f(n) : 0;
}
int LZ4_compress_fast(const char* a, char* b, int c, int d, int e)
{
	auto f = reinterpret_cast<int (*)(const char*, char*, int, int, int)>(lz4_sym("LZ4_compress_fast"));
	return f ? f(a, b, c, d, e) : -1;
}
int LZ4_decompress_safe(const char* a, char* b, int c, int d)
{
	auto f = reinterpret_cast<int (*)(const char*, char*, int, int)>(lz4_sym("LZ4_decompress_safe"));
	return f ? f(a, b, c, d) : -1;
}
int LZ4_compress_HC(const char* a, char* b, int c, int d, int e)
{
	auto f = reinterpret_cast<int (*)(const char*, char*, int, int, int)>(lz4_sym("LZ4_compress_HC"));
	return f ? f(a, b, c, d, e) : -1;
}
}

namespace
{
using ufo::map::Code;
using ufo::map::CodeMap;
using ufo::map::Color;
using ufo::map::DepthType;
using ufo::map::Key;
using ufo::map::Point3;
using ufo::map::Point3Color;
using ufo::map::PointCloud;
using ufo::map::PointCloudColor;

struct NodeRec {
	uint64_t code;
	uint32_t depth;
	float occ;
	uint8_t rgb[3];
	uint8_t flags;  // bit0 contains_free, bit1 contains_unknown, bit2 is_leaf
};

inline void colorOf(ufo::map::OccupancyNode<float> const&, uint8_t* rgb)
{
	rgb[0] = rgb[1] = rgb[2] = 0;
}
inline void colorOf(ufo::map::ColorOccupancyNode<float> const& v, uint8_t* rgb)
{
	rgb[0] = v.color.r;
	rgb[1] = v.color.g;
	rgb[2] = v.color.b;
}

template <class MAP>
class Probe : public MAP
{
 public:
	using MAP::MAP;

	// Pre-order walk of the real pointer tree.  leaves=true: emit nodes without
	// children (depth-0 voxels and collapsed / never-split inner nodes);
	// leaves=false: emit nodes that have children.
	void walk(std::vector<NodeRec>& out, bool leaves) const
	{
		walkRec(this->getRoot(), this->getTreeDepthLevels(), 0, out, leaves);
	}

	// Value field in code order: every depth-0 voxel whose payload is not the default
	// (0.0, black), collapsed / never-split nodes expanded.  Pre-order over the children in
	// index order is ascending Morton order, so the output is sorted.  Pass null arrays to count.
	size_t field(uint64_t* codes, float* occ, uint8_t* rgb, size_t cap) const
	{
		size_t n = 0;
		fieldRec(this->getRoot(), this->getTreeDepthLevels(), 0, codes, occ, rgb, cap, n);
		return n;
	}

	// getNodePath semantics (octree.h:957-972): deepest existing node on the
	// path to `code`.
	bool nodeAt(uint64_t code, unsigned depth, NodeRec& rec)
	{
		auto [path, d] = this->getNodePath(Code(code, depth));
		rec.code = code;
		rec.depth = d;
		rec.occ = path[d]->value.occupancy;
		colorOf(path[d]->value, rec.rgb);
		rec.flags = 0;
		if (d > 0) {
			auto const& in = static_cast<typename MAP::INNER_NODE const&>(*path[d]);
			rec.flags = (in.contains_free ? 1 : 0) | (in.contains_unknown ? 2 : 0) |
			            (in.is_leaf ? 4 : 0);
		} else {
			rec.flags = (this->isFree(*path[0]) ? 1 : 0) | (this->isUnknown(*path[0]) ? 2 : 0) | 4;
		}
		return d == depth;
	}

	// Protected freeSpace (occupancy_map_base.h:1229-1259) exposed: returns the
	// per-scan free set exactly as the integrator builds it.
	void freeSet(Point3 const& origin, PointCloud const& ends, unsigned depth, bool simple,
	             unsigned early_stopping, std::vector<uint64_t>& codes) const
	{
		CodeMap<float> free_hits;
		this->freeSpace(origin, ends, free_hits, -1.0f, depth, simple, early_stopping);
		codes.clear();
		codes.reserve(free_hits.size());
		for (auto const& [code, value] : free_hits) {
			codes.push_back(code.getCode());
		}
	}

	bool moveLine(Point3& a, Point3& b) const { return this->moveLineInside(a, b); }

	// The stored (double) log-odds parameters (occupancy_map_base.h:1537-1542).
	void sensorModel(double* out6) const
	{
		out6[0] = this->occupied_thres_log_;
		out6[1] = this->free_thres_log_;
		out6[2] = this->prob_hit_log_;
		out6[3] = this->prob_miss_log_;
		out6[4] = this->clamping_thres_min_log_;
		out6[5] = this->clamping_thres_max_log_;
	}

 private:
	static void emit(uint64_t code, unsigned depth, float o, uint8_t const* c, uint64_t* codes, float* occ,
	                 uint8_t* rgb, size_t cap, size_t& n)
	{
		if (o == 0.0f && !(c[0] | c[1] | c[2])) return;
		const uint64_t count = uint64_t(1) << (3 * depth);
		if (codes && n + count <= cap) {
			for (uint64_t i = 0; i < count; ++i) {
				codes[n + i] = code + i;
				occ[n + i] = o;
				if (rgb) std::memcpy(rgb + 3 * (n + i), c, 3);
			}
		}
		n += count;
	}

	template <class NODE>
	void fieldRec(NODE const& node, unsigned depth, uint64_t code, uint64_t* codes, float* occ, uint8_t* rgb,
	              size_t cap, size_t& n) const
	{
		if (node.is_leaf) {
			uint8_t c[3];
			colorOf(node.value, c);
			emit(code, depth, node.value.occupancy, c, codes, occ, rgb, cap, n);
			return;
		}
		unsigned cd = depth - 1;
		for (unsigned i = 0; i < 8; ++i) {
			uint64_t ccode = code + (uint64_t(i) << (3 * cd));
			if (0 == cd) {
				auto const& leaf = MAP::getLeafChild(node, i);
				uint8_t c[3];
				colorOf(leaf.value, c);
				emit(ccode, 0, leaf.value.occupancy, c, codes, occ, rgb, cap, n);
			} else {
				fieldRec(MAP::getInnerChild(node, i), cd, ccode, codes, occ, rgb, cap, n);
			}
		}
	}

	template <class NODE>
	void walkRec(NODE const& node, unsigned depth, uint64_t code, std::vector<NodeRec>& out,
	             bool leaves) const
	{
		bool is_leaf = node.is_leaf;
		if (is_leaf == leaves) {
			NodeRec r;
			r.code = code;
			r.depth = depth;
			r.occ = node.value.occupancy;
			colorOf(node.value, r.rgb);
			r.flags = (node.contains_free ? 1 : 0) | (node.contains_unknown ? 2 : 0) |
			          (is_leaf ? 4 : 0);
			out.push_back(r);
		}
		if (is_leaf) {
			return;
		}
		unsigned cd = depth - 1;
		for (unsigned i = 0; i < 8; ++i) {
			uint64_t ccode = code + (uint64_t(i) << (3 * cd));
			if (0 == cd) {
				if (leaves) {
					auto const& leaf = MAP::getLeafChild(node, i);
					NodeRec r;
					r.code = ccode;
					r.depth = 0;
					r.occ = leaf.value.occupancy;
					colorOf(leaf.value, r.rgb);
					r.flags = (this->isFree(leaf) ? 1 : 0) | (this->isUnknown(leaf) ? 2 : 0) | 4;
					out.push_back(r);
				}
			} else {
				walkRec(MAP::getInnerChild(node, i), cd, ccode, out, leaves);
			}
		}
	}
};

struct RefMap {
	bool color;
	Probe<ufo::map::OccupancyMap>* mono = nullptr;
	Probe<ufo::map::OccupancyMapColor>* col = nullptr;
	std::vector<NodeRec> scratch;
	~RefMap()
	{
		delete mono;
		delete col;
	}
};

template <class F>
auto withMap(RefMap* m, F&& f)
{
	if (m->color) {
		return f(*m->col);
	}
	return f(*m->mono);
}
}  // namespace

extern "C" {

void* ufo_ref_create(double resolution, unsigned depth_levels, int automatic_pruning,
                     double occupied_thres, double free_thres, double prob_hit,
                     double prob_miss, double clamp_min, double clamp_max, int color)
{
	try {
		RefMap* m = new RefMap;
		m->color = color != 0;
		if (m->color) {
			m->col = new Probe<ufo::map::OccupancyMapColor>(
			    resolution, depth_levels, automatic_pruning != 0, occupied_thres, free_thres,
			    prob_hit, prob_miss, clamp_min, clamp_max);
			m->col->enableMinMaxChangeDetection(true);
		} else {
			m->mono = new Probe<ufo::map::OccupancyMap>(
			    resolution, depth_levels, automatic_pruning != 0, occupied_thres, free_thres,
			    prob_hit, prob_miss, clamp_min, clamp_max);
			m->mono->enableMinMaxChangeDetection(true);
		}
		return m;
	} catch (std::exception const&) {
		return nullptr;
	}
}

void ufo_ref_destroy(void* h) { delete static_cast<RefMap*>(h); }

// xyz: n*3 doubles; rgb: n*3 bytes or NULL.  discrete!=0 selects
// insertPointCloudDiscrete.  Returns wall seconds spent inside the reference call
// (steady_clock, as ufomap_mapping/src/server.cpp:111-125 times it).
double ufo_ref_insert(void* h, const double* origin, const double* xyz, const uint8_t* rgb,
                      size_t n, double max_range, unsigned depth, int simple,
                      unsigned early_stopping, int discrete, int async)
{
	RefMap* m = static_cast<RefMap*>(h);
	Point3 o(origin[0], origin[1], origin[2]);
	double secs = 0;
	if (rgb) {
		PointCloudColor cloud;
		cloud.reserve(n);
		for (size_t i = 0; i < n; ++i) {
			cloud.push_back(Point3Color(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], rgb[3 * i],
			                            rgb[3 * i + 1], rgb[3 * i + 2]));
		}
		auto t0 = std::chrono::steady_clock::now();
		if (m->color) {
			// OccupancyMapColor::insertPointCloud(PointCloudColor) does not compile in
			// the reference (occupancy_map_color.h:113); only the discrete variant exists.
			if (!discrete) {
				return -1.0;
			}
			m->col->insertPointCloudDiscrete(o, cloud, max_range, depth, simple != 0,
			                                 early_stopping, async != 0);
			m->col->insertPointCloudWait();
		} else {
			if (!discrete) {
				return -1.0;
			}
			m->mono->insertPointCloudDiscrete(o, cloud, max_range, depth, simple != 0,
			                                  early_stopping, async != 0);
			m->mono->insertPointCloudWait();
		}
		secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
	} else {
		PointCloud cloud;
		cloud.reserve(n);
		for (size_t i = 0; i < n; ++i) {
			cloud.push_back(Point3(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]));
		}
		auto t0 = std::chrono::steady_clock::now();
		withMap(m, [&](auto& map) {
			if (discrete) {
				map.insertPointCloudDiscrete(o, cloud, max_range, depth, simple != 0,
				                             early_stopping, async != 0);
			} else {
				map.insertPointCloud(o, cloud, max_range, depth, simple != 0, early_stopping,
				                     async != 0);
			}
			map.insertPointCloudWait();
		});
		secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
	}
	return secs;
}

// Two-phase dump: ufo_ref_walk() fills an internal buffer and returns the count;
// ufo_ref_walk_fetch() copies it out as SoA.
size_t ufo_ref_walk(void* h, int leaves)
{
	RefMap* m = static_cast<RefMap*>(h);
	m->scratch.clear();
	withMap(m, [&](auto& map) { map.walk(m->scratch, leaves != 0); });
	return m->scratch.size();
}

void ufo_ref_walk_fetch(void* h, uint64_t* codes, uint32_t* depths, float* occ, uint8_t* rgb,
                        uint8_t* flags)
{
	RefMap* m = static_cast<RefMap*>(h);
	for (size_t i = 0; i < m->scratch.size(); ++i) {
		NodeRec const& r = m->scratch[i];
		if (codes) codes[i] = r.code;
		if (depths) depths[i] = r.depth;
		if (occ) occ[i] = r.occ;
		if (rgb) std::memcpy(rgb + 3 * i, r.rgb, 3);
		if (flags) flags[i] = r.flags;
	}
	m->scratch.clear();
	m->scratch.shrink_to_fit();
}

// Sorted, expanded value field (see Probe::field); null arrays: count only.
size_t ufo_ref_field(void* h, uint64_t* codes, float* occ, uint8_t* rgb, size_t cap)
{
	RefMap* m = static_cast<RefMap*>(h);
	return withMap(m, [&](auto& map) { return map.field(codes, occ, rgb, cap); });
}

// Batched ufo_ref_node: deepest existing node on the path to every (code, depth).
void ufo_ref_node_batch(void* h, const uint64_t* codes, const uint32_t* depths, size_t n, float* occ,
                        uint8_t* rgb, uint8_t* flags, uint32_t* found_depth)
{
	RefMap* m = static_cast<RefMap*>(h);
	withMap(m, [&](auto& map) {
		for (size_t i = 0; i < n; ++i) {
			NodeRec r;
			map.nodeAt(codes[i], depths[i], r);
			occ[i] = r.occ;
			if (rgb) std::memcpy(rgb + 3 * i, r.rgb, 3);
			if (flags) flags[i] = r.flags;
			if (found_depth) found_depth[i] = r.depth;
		}
	});
}

// Change detection (occupancy_map_base.h:779-790): the reference's own changes_ set.
void ufo_ref_enable_changes(void* h, int enable)
{
	RefMap* m = static_cast<RefMap*>(h);
	withMap(m, [&](auto& map) { map.enableChangeDetection(enable != 0); });
}

void ufo_ref_reset_changes(void* h)
{
	RefMap* m = static_cast<RefMap*>(h);
	withMap(m, [&](auto& map) { map.resetChangeDetection(); });
}

size_t ufo_ref_changes(void* h, uint64_t* codes, uint32_t* depths, size_t cap)
{
	RefMap* m = static_cast<RefMap*>(h);
	return withMap(m, [&](auto& map) {
		size_t n = 0;
		for (auto it = map.changesBegin(); it != map.changesEnd(); ++it, ++n) {
			if (codes && n < cap) {
				codes[n] = (*it).getCode();
				depths[n] = (*it).getDepth();
			}
		}
		return n;
	});
}

// beginLeaves(occupied, free, unknown, contains = false, min_depth) with an optional AABB
// (occupancy_map_base.h:130-216): the reference's own leaf iteration, for the filtered read-out.
size_t ufo_ref_leaves(void* h, int occupied, int free_space, int unknown, const double* box6, unsigned min_depth,
                      uint64_t* codes, uint32_t* depths, float* occ, size_t cap)
{
	RefMap* m = static_cast<RefMap*>(h);
	return withMap(m, [&](auto& map) {
		size_t n = 0;
		auto emit = [&](auto it, auto end) {
			for (; it != end; ++it, ++n) {
				if (codes && n < cap) {
					codes[n] = it.getCode().getCode();
					depths[n] = it.getDepth();
					occ[n] = (float)it->occupancy;
				}
			}
		};
		if (box6) {
			ufo::geometry::AABB box(Point3(box6[0], box6[1], box6[2]), Point3(box6[3], box6[4], box6[5]));
			emit(map.beginLeaves(box, occupied != 0, free_space != 0, unknown != 0, false, min_depth), map.endLeaves());
		} else {
			emit(map.beginLeaves(occupied != 0, free_space != 0, unknown != 0, false, min_depth), map.endLeaves());
		}
		return n;
	});
}

// Returns 1 if a node exists at exactly (code, depth); out describes the deepest
// existing node on the path either way.
int ufo_ref_node(void* h, uint64_t code, unsigned depth, float* occ, uint8_t* rgb,
                 uint8_t* flags, unsigned* found_depth)
{
	RefMap* m = static_cast<RefMap*>(h);
	NodeRec r;
	bool exact = withMap(m, [&](auto& map) { return map.nodeAt(code, depth, r); });
	*occ = r.occ;
	std::memcpy(rgb, r.rgb, 3);
	*flags = r.flags;
	*found_depth = r.depth;
	return exact ? 1 : 0;
}

size_t ufo_ref_compute_ray(void* h, const double* origin, const double* end, double max_range,
                           unsigned depth, uint64_t* codes, size_t cap)
{
	RefMap* m = static_cast<RefMap*>(h);
	auto ray = withMap(m, [&](auto& map) {
		return map.computeRay(Point3(origin[0], origin[1], origin[2]),
		                      Point3(end[0], end[1], end[2]), max_range, depth);
	});
	for (size_t i = 0; i < ray.size() && i < cap; ++i) {
		codes[i] = ray[i].getCode();
	}
	return ray.size();
}

// The integrator's free set (protected freeSpace) for rays origin -> ends[i].
size_t ufo_ref_free_set(void* h, const double* origin, const double* ends, size_t n,
                        unsigned depth, int simple, unsigned early_stopping, uint64_t* codes,
                        size_t cap)
{
	RefMap* m = static_cast<RefMap*>(h);
	PointCloud cloud;
	cloud.reserve(n);
	for (size_t i = 0; i < n; ++i) {
		cloud.push_back(Point3(ends[3 * i], ends[3 * i + 1], ends[3 * i + 2]));
	}
	std::vector<uint64_t> out;
	withMap(m, [&](auto& map) {
		map.freeSet(Point3(origin[0], origin[1], origin[2]), cloud, depth, simple != 0,
		            early_stopping, out);
	});
	for (size_t i = 0; i < out.size() && i < cap; ++i) {
		codes[i] = out[i];
	}
	return out.size();
}

void ufo_ref_to_key(void* h, const double* xyz, unsigned depth, uint32_t* key)
{
	RefMap* m = static_cast<RefMap*>(h);
	Key k = withMap(m, [&](auto& map) { return map.toKey(Point3(xyz[0], xyz[1], xyz[2]), depth); });
	key[0] = k[0];
	key[1] = k[1];
	key[2] = k[2];
}

uint64_t ufo_ref_to_code(void* h, const double* xyz, unsigned depth)
{
	RefMap* m = static_cast<RefMap*>(h);
	return withMap(m, [&](auto& map) {
		return map.toCode(Point3(xyz[0], xyz[1], xyz[2]), depth).getCode();
	});
}

uint64_t ufo_ref_key_to_code(const uint32_t* key, unsigned depth)
{
	return Code(Key(key[0], key[1], key[2], depth)).getCode();
}

void ufo_ref_code_to_key(uint64_t code, unsigned depth, uint32_t* key)
{
	Key k = Code(code, depth).toKey();
	key[0] = k[0];
	key[1] = k[1];
	key[2] = k[2];
}

void ufo_ref_key_to_coord(void* h, const uint32_t* key, unsigned depth, double* xyz)
{
	RefMap* m = static_cast<RefMap*>(h);
	Point3 p = withMap(m, [&](auto& map) { return map.toCoord(Key(key[0], key[1], key[2], depth)); });
	xyz[0] = p[0];
	xyz[1] = p[1];
	xyz[2] = p[2];
}

int ufo_ref_move_line_inside(void* h, double* a, double* b)
{
	RefMap* m = static_cast<RefMap*>(h);
	Point3 pa(a[0], a[1], a[2]), pb(b[0], b[1], b[2]);
	bool ok = withMap(m, [&](auto& map) { return map.moveLine(pa, pb); });
	for (int i = 0; i < 3; ++i) {
		a[i] = pa[i];
		b[i] = pb[i];
	}
	return ok ? 1 : 0;
}

void ufo_ref_change_bbox(void* h, double* mn, double* mx)
{
	RefMap* m = static_cast<RefMap*>(h);
	withMap(m, [&](auto& map) {
		for (int i = 0; i < 3; ++i) {
			mn[i] = map.minChange()[i];
			mx[i] = map.maxChange()[i];
		}
	});
}

void ufo_ref_reset_change_bbox(void* h)
{
	RefMap* m = static_cast<RefMap*>(h);
	withMap(m, [&](auto& map) { map.resetMinMaxChangeDetection(); });
}

// Sensor model constants as the reference stores them (double log-odds).
void ufo_ref_sensor_model(void* h, double* out6)
{
	RefMap* m = static_cast<RefMap*>(h);
	withMap(m, [&](auto& map) { map.sensorModel(out6); });
}

size_t ufo_ref_memory_usage(void* h)
{
	RefMap* m = static_cast<RefMap*>(h);
	return withMap(m, [&](auto& map) { return size_t(map.memoryUsage()); });
}

// Octree::write(std::ostream&) with compress = false (octree.h:833-864): the complete file image.
// Returns its size; copies it when it fits into cap.
size_t ufo_ref_write(void* h, uint8_t* buf, size_t cap)
{
	RefMap* m = static_cast<RefMap*>(h);
	std::stringstream ss(std::ios_base::in | std::ios_base::out | std::ios_base::binary);
	withMap(m, [&](auto& map) { return map.write(ss, false); });
	const std::string str = ss.str();
	if (buf && str.size() <= cap) std::memcpy(buf, str.data(), str.size());
	return str.size();
}

// Octree::write(ostream, compress = true, min_depth, acceleration, level): LZ4-compressed file image
size_t ufo_ref_write_compressed(void* h, unsigned min_depth, int acceleration, int level, uint8_t* buf, size_t cap)
{
	RefMap* m = static_cast<RefMap*>(h);
	std::stringstream ss(std::ios_base::in | std::ios_base::out | std::ios_base::binary);
	bool ok = withMap(m, [&](auto& map) { return map.write(ss, true, min_depth, acceleration, level); });
	if (!ok) return 0;
	const std::string str = ss.str();
	if (buf && str.size() <= cap) std::memcpy(buf, str.data(), str.size());
	return str.size();
}

// setValueVolume(AABB(min, max), occupancy probability, min_depth) (occupancy_map_base.h:492-518),
// the call behind the server's robot clearing and clear_volume service (server.cpp:152-154, :354).
void ufo_ref_set_value_volume(void* h, const double* box6, double occupancy, unsigned min_depth)
{
	RefMap* m = static_cast<RefMap*>(h);
	ufo::geometry::AABB aabb(ufo::geometry::Point(box6[0], box6[1], box6[2]),
	                         ufo::geometry::Point(box6[3], box6[4], box6[5]));
	withMap(m, [&](auto& map) { map.setValueVolume(aabb, occupancy, min_depth); });
}

// Octree::writeData(stream, bounding_volume, compress = false, min_depth) (octree.h:885-917) -- the
// payload ufoToMsg puts into a UFOMap message (ufomap_msgs/conversions.h:161-185).  box6 = AABB
// min xyz, max xyz (the server's change box, server.cpp:184) or NULL for the whole map.
size_t ufo_ref_write_data(void* h, const double* box6, unsigned min_depth, uint8_t* buf, size_t cap)
{
	RefMap* m = static_cast<RefMap*>(h);
	std::stringstream ss(std::ios_base::in | std::ios_base::out | std::ios_base::binary);
	ufo::geometry::BoundingVolume bv;
	if (box6) {
		bv.add(ufo::geometry::AABB(ufo::geometry::Point(box6[0], box6[1], box6[2]),
		                           ufo::geometry::Point(box6[3], box6[4], box6[5])));
	}
	withMap(m, [&](auto& map) { return map.writeData(ss, bv, false, min_depth); });
	const std::string str = ss.str();
	if (buf && str.size() <= cap) std::memcpy(buf, str.data(), str.size());
	return str.size();
}

// Octree::write(std::ostream&, bounding_volume, compress = false, min_depth) (octree.h:826-864):
// file image of a region -- the server's save_map service (server.cpp:381-392).
size_t ufo_ref_write_region(void* h, const double* box6, unsigned min_depth, uint8_t* buf, size_t cap)
{
	RefMap* m = static_cast<RefMap*>(h);
	std::stringstream ss(std::ios_base::in | std::ios_base::out | std::ios_base::binary);
	ufo::geometry::BoundingVolume bv;
	if (box6) {
		bv.add(ufo::geometry::AABB(ufo::geometry::Point(box6[0], box6[1], box6[2]),
		                           ufo::geometry::Point(box6[3], box6[4], box6[5])));
	}
	withMap(m, [&](auto& map) { return map.write(ss, bv, false, min_depth); });
	const std::string str = ss.str();
	if (buf && str.size() <= cap) std::memcpy(buf, str.data(), str.size());
	return str.size();
}

// Octree::clear(resolution, depth_levels) (octree.h:541-560)
void ufo_ref_clear(void* h, double resolution, unsigned depth_levels)
{
	RefMap* m = static_cast<RefMap*>(h);
	withMap(m, [&](auto& map) { map.clear(resolution, depth_levels); });
}

// Octree::readData(stream, bounding_volume, resolution, depth_levels) (octree.h:735-774): merge a
// node stream (a UFOMap message, msgToUfo ufomap_msgs/conversions.h:122-134) into the map.
int ufo_ref_read_data(void* h, const double* box6, const uint8_t* buf, size_t size)
{
	RefMap* m = static_cast<RefMap*>(h);
	std::stringstream ss(std::string(reinterpret_cast<const char*>(buf), size),
	                     std::ios_base::in | std::ios_base::out | std::ios_base::binary);
	ufo::geometry::BoundingVolume bv;
	if (box6) {
		bv.add(ufo::geometry::AABB(ufo::geometry::Point(box6[0], box6[1], box6[2]),
		                           ufo::geometry::Point(box6[3], box6[4], box6[5])));
	}
	return withMap(m, [&](auto& map) {
		return map.readData(ss, bv, map.getResolution(), map.getTreeDepthLevels());
	}) ? 1 : 0;
}

// Octree::read(std::istream&) (octree.h:699-733): replaces the map's content.  1 = ok.
int ufo_ref_read(void* h, const uint8_t* buf, size_t size)
{
	RefMap* m = static_cast<RefMap*>(h);
	std::stringstream ss(std::string(reinterpret_cast<const char*>(buf), size),
	                     std::ios_base::in | std::ios_base::out | std::ios_base::binary);
	return withMap(m, [&](auto& map) { return map.read(ss); }) ? 1 : 0;
}

// The cloud transform of the insertPointCloud(..., frame_origin, ...) overloads, through the
// reference's own PointCloud::transform (point_cloud.h:157-166).  pose7 = tx ty tz qw qx qy qz.
void ufo_ref_transform(const double* pose7, const double* xyz, size_t n, double* out)
{
	ufo::math::Pose6 pose(pose7[0], pose7[1], pose7[2], pose7[3], pose7[4], pose7[5], pose7[6]);
	PointCloud cloud;
	cloud.reserve(n);
	for (size_t i = 0; i < n; ++i) cloud.push_back(Point3(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]));
	cloud.transform(pose);
	for (size_t i = 0; i < n; ++i) {
		out[3 * i + 0] = cloud[i][0];
		out[3 * i + 1] = cloud[i][1];
		out[3 * i + 2] = cloud[i][2];
	}
}

// Pose6(x, y, z, roll, pitch, yaw) (math/pose6.h:71-74) as tx ty tz qw qx qy qz.
void ufo_ref_pose_from_rpy(double x, double y, double z, double roll, double pitch, double yaw,
                           double* pose7)
{
	ufo::math::Pose6 pose(x, y, z, roll, pitch, yaw);
	pose7[0] = pose.translation()[0];
	pose7[1] = pose.translation()[1];
	pose7[2] = pose.translation()[2];
	pose7[3] = pose.rotation().w();
	pose7[4] = pose.rotation().x();
	pose7[5] = pose.rotation().y();
	pose7[6] = pose.rotation().z();
}

}  // extern "C"
