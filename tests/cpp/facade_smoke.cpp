// Reads like a program written against the reference, include lines included (include/ufo/ holds
// forwarding headers under the reference's own paths).  Built and run by tests/test_facade.py.
#include <ufo/map/occupancy_map.h>
#include <ufo/map/occupancy_map_color.h>

#include <cstdio>
#include <cstring>
#include <sstream>

int main(int argc, char** argv)
{
	using namespace ufo::map;
	bool compile_only = argc > 1 && 0 == std::strcmp(argv[1], "--no-device");
	// Pose6 / PointCloud::transform are host code (math/pose6.h:115-125, point_cloud.h:157-166)
	ufo::math::Pose6 pose(1.0, 2.0, 3.0, 0.0, 0.0, 1.5707963267948966);
	PointCloud one;
	one.push_back(Point3(1.0, 0.0, 0.0));
	one.transform(pose);
	if (std::fabs(one[0].x() - 1.0) > 1e-12 || std::fabs(one[0].y() - 3.0) > 1e-12 || one[0].z() != 3.0) return 10;
	if (compile_only) {
		// constructor arguments are validated before any device is touched (octree.h:931-935)
		try {
			OccupancyMap bad(0.1, 40);
			return 2;
		} catch (std::invalid_argument const&) {
		}
		std::puts("facade ok (no device)");
		return 0;
	}

	OccupancyMap map(0.02);
	PointCloud cloud;
	cloud.push_back(Point3(1.0, 0.0, 0.0));
	map.insertPointCloud(Point3(0, 0, 0), cloud, 5.0);

	// end voxel = float(hit) + float(miss), mid voxel = float(miss), origin untouched
	double end = map.getOccupancy(Point3(1.0, 0, 0)), mid = map.getOccupancy(Point3(0.5, 0, 0));
	if (!map.isOccupied(Point3(1.0, 0, 0)) || !map.isFree(Point3(0.5, 0, 0)) || !map.isUnknown(Point3(0, 0, 0))) return 3;
	if (!(end > 0.6 && end < 0.62) || !(mid > 0.39 && mid < 0.41)) return 4;

	CodeRay ray = map.computeRay(Point3(0, 0, 0), Point3(0.1, 0.1, 0.1));
	if (ray.size() != 15) return 5;
	if (map.toCode(Point3(1.0, 0, 0)).toKey() != map.toKey(Point3(1.0, 0, 0))) return 6;

	OccupancyMapColor cmap(0.02);
	PointCloudColor ccloud;
	ccloud.push_back(Point3Color(0.5, 0.1, 0.2, 200, 100, 50));
	cmap.insertPointCloudDiscrete(Point3(0, 0, 0), ccloud, -1, 0, false, 0, true);
	cmap.insertPointCloudWait();
	if (!cmap.insertPointCloudDone()) return 7;
	Color c = cmap.getColor(Point3(0.5, 0.1, 0.2));
	if (c != Color(200, 100, 50)) return 8;

	// early_stopping is rejected, not silently ignored
	map.insertPointCloud(Point3(0, 0, 0), cloud, 5.0, 0, false, 3);
	if (map.lastStatus() != UFO_B200_E_UNSUPPORTED) return 9;

	// frame overload (device-side transform) == host transform + plain insert (occupancy_map_base.h:313-327)
	ufo::math::Pose6 tilt(0.4, -0.2, 0.1, 0.03, -0.02, 0.9);
	PointCloud local;
	for (int i = 0; i < 50; ++i) local.push_back(Point3(2.0 + 0.07 * i, 0.031 * i, 0.011 * i - 0.2));
	PointCloud world = local;
	world.transform(tilt);
	OccupancyMap fused(0.05), staged(0.05);
	fused.insertPointCloud(tilt.translation(), local, tilt, 10.0);
	staged.insertPointCloud(tilt.translation(), world, 10.0);
	for (std::size_t i = 0; i < world.size(); ++i) {
		if (fused.getOccupancy(world[i]) != staged.getOccupancy(world[i]) || !fused.isOccupied(world[i])) return 11;
	}
	// Octree::write(std::ostream&) (octree.h:812-864)
	std::ostringstream os;
	if (!fused.write(os) || os.str().compare(0, 13, "# UFOMap file") != 0) return 12;
	if (os.str().find("id occupancy_map\nresolution 0.05\ndepth_levels 16\ncompressed 0\n") == std::string::npos) return 13;
	// compress = true: one LZ4 block behind a header that says "compressed 1" (octree.h:1428-1456)
	std::ostringstream zs;
	if (!fused.write(zs, /*compress*/ true) || zs.str().find("compressed 1\n") == std::string::npos ||
	    zs.str().size() >= os.str().size())
		return 14;
	// Octree::writeData with the change box, as ufoToMsg calls it (ufomap_msgs/conversions.h:176-178)
	std::ostringstream part, whole;
	ufo::geometry::AABB aabb(world[10] - Point3(0.3, 0.3, 0.3), world[10] + Point3(0.3, 0.3, 0.3));
	int n_part = fused.writeData(part, aabb, false, 2), n_whole = fused.writeData(whole, false, 0);
	if (n_part <= 0 || n_whole <= n_part || (std::size_t)n_whole != whole.str().size()) return 15;
	// clear the robot's volume like the server does (server.cpp:150-154)
	Point3 r(0.2, 0.2, 0.1);
	fused.setValueVolume(ufo::geometry::AABB(world[20] - r, world[20] + r), fused.getClampingThresMin(), 0);
	if (fused.lastStatus() != UFO_B200_OK || !fused.isFree(world[20])) return 16;
	// readData: the node stream of one map merged into an empty one gives the same occupancy
	{
		std::stringstream data(std::ios_base::in | std::ios_base::out | std::ios_base::binary);
		int n_data = fused.writeData(data, false, 0);
		OccupancyMap copy(0.05);
		if (n_data <= 0 || !copy.readData(data, 0.05, 16, n_data, false)) return 24;
		for (std::size_t i = 0; i < world.size(); ++i)
			if (copy.getOccupancy(world[i]) != fused.getOccupancy(world[i])) return 25;
		// castRay from the sensor along a measured ray hits the ray's occupied end voxel
		auto hit = copy.castRay(tilt.translation() + (world[30] - tilt.translation()) * 0.05, world[30] - tilt.translation(), true, 20.0);
		if (!hit || !copy.isOccupied(*hit)) return 26;
	}
	// a setter changes ONE stored log-odds; the others stay bit for bit (occupancy_map_base.h:748-773)
	{
		OccupancyMap a(0.05), b(0.05);
		double before[6], after[6];
		ufo_b200_sensor_model_logit(a.handle(), before);
		a.setProbMiss(0.4);  // same value again: nothing may move
		ufo_b200_sensor_model_logit(a.handle(), after);
		for (int i = 0; i < 6; ++i)
			if (std::memcmp(&before[i], &after[i], sizeof(double)) != 0) return 17;
		a.setProbMiss(0.35);
		ufo_b200_sensor_model_logit(a.handle(), after);
		for (int i = 0; i < 6; ++i)
			if ((i == 3) == (std::memcmp(&before[i], &after[i], sizeof(double)) == 0)) return 18;
		// hits are still applied with the untouched hit increment: same value as on a fresh map
		a.insertPointCloud(tilt.translation(), world, 10.0);
		b.insertPointCloud(tilt.translation(), world, 10.0);
		a.insertPointCloud(tilt.translation(), world, 10.0);
		b.insertPointCloud(tilt.translation(), world, 10.0);
		if (a.getOccupancy(world[7]) == b.getOccupancy(world[7])) return 19;  // different miss, end voxel gets both
	}
	// Octree::clear(resolution, depth_levels): the geometry getters follow (server.cpp:364-378)
	fused.clear(0.1, 14);
	if (fused.getResolution() != 0.1 || fused.getTreeDepthLevels() != 14) return 20;
	if (fused.getMax().x() != 0.1 * 8192 || fused.getMin().x() != -0.1 * 8192) return 21;
	if (!fused.isInside(Point3(800.0, 0, 0)) || fused.isInside(Point3(900.0, 0, 0))) return 22;
	fused.insertPointCloud(tilt.translation(), world, 10.0);
	if (!fused.isOccupied(world[3])) return 23;
	std::puts("facade ok");
	return 0;
}
