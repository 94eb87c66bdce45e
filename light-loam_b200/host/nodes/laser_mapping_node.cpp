// alaserMapping with process() LM:1581-2168 replaced by ll_mapping_step.  Same subscriptions (LM:2371-2377), the
// synchronisation / frame dropping of LM:1511-1575, and /aft_mapped_to_init (frame "rslidar", child "/aft_mapped",
// LM:2262-2274), /aft_mapped_path, tf rslidar -> /aft_mapped and the RESULT_PATH trajectory line (LM:2284-2325).
// (/laser_cloud_surround, /laser_cloud_map and /velodyne_cloud_registered are visualisation outputs of the map and are
// not reproduced: the map lives in HBM.)
#ifdef LL_WITH_ROS
#include <nav_msgs/Odometry.h>
#include <nav_msgs/Path.h>
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#include <pcl_conversions/pcl_conversions.h>
#include <ros/ros.h>
#include <sensor_msgs/PointCloud2.h>
#include <tf/transform_broadcaster.h>

#include <Eigen/Dense>
#include <fstream>
#include <mutex>
#include <queue>
#include <thread>

#include "lightloam_b200.h"

static std::queue<sensor_msgs::PointCloud2ConstPtr> qCorner, qSurf, qFull;
static std::queue<nav_msgs::Odometry::ConstPtr> qOdom;
static std::mutex mBuf;
static ll_ctx* g_ll = nullptr;
static std::string RESULT_PATH;
static ros::Publisher pubOdomAftMapped, pubPath;

// a PointXYZI payload read in place: point_step 32, x,y,z at 0, intensity at byte 16 (LM:1551-1563 fromROSMsg)
static ll_cloud_view view(const sensor_msgs::PointCloud2& m)
{
    return ll_cloud_view{reinterpret_cast<const float*>(m.data.data()), (int)(m.width * m.height), (int)m.point_step};
}

static void process()
{
    nav_msgs::Path path;
    bool init_flag = true;
    Eigen::Matrix4f H_init = Eigen::Matrix4f::Identity();
    while (ros::ok()) {
        while (true) {
            std::unique_lock<std::mutex> lk(mBuf);
            if (qCorner.empty() || qSurf.empty() || qFull.empty() || qOdom.empty()) break;
            const double tc = qCorner.front()->header.stamp.toSec();  // LM:1511-1548
            while (!qOdom.empty() && qOdom.front()->header.stamp.toSec() < tc) qOdom.pop();
            while (!qSurf.empty() && qSurf.front()->header.stamp.toSec() < tc) qSurf.pop();
            while (!qFull.empty() && qFull.front()->header.stamp.toSec() < tc) qFull.pop();
            if (qOdom.empty() || qSurf.empty() || qFull.empty()) break;
            if (qSurf.front()->header.stamp.toSec() != tc || qFull.front()->header.stamp.toSec() != tc || qOdom.front()->header.stamp.toSec() != tc) break;
            const sensor_msgs::PointCloud2ConstPtr mCorner = qCorner.front(), mSurf = qSurf.front();
            const nav_msgs::Odometry odom = *qOdom.front();
            qCorner.pop(); qSurf.pop(); qFull.pop(); qOdom.pop();
            while (!qCorner.empty()) qCorner.pop();  // LM:1571-1575 real-time frame dropping
            lk.unlock();
            const double qi[4] = {odom.pose.pose.orientation.x, odom.pose.pose.orientation.y, odom.pose.pose.orientation.z, odom.pose.pose.orientation.w};
            const double ti[3] = {odom.pose.pose.position.x, odom.pose.pose.position.y, odom.pose.pose.position.z};
            double q[4], t[3];
            const ll_cloud_view vc = view(*mCorner), vs = view(*mSurf);
            const int rc = ll_mapping_step(g_ll, vc, vs, qi, ti, q, t);
            if (rc == LL_W_FEW_CORRESPONDENCES) ROS_WARN("time Map corner and surf num are not enough");  // LM:2097-2100
            if (rc < 0) { ROS_WARN("lightloam_b200: %s", ll_strerror(rc)); continue; }
            nav_msgs::Odometry out;
            out.header.frame_id = "rslidar";
            out.child_frame_id = "/aft_mapped";
            out.header.stamp = odom.header.stamp;
            out.pose.pose.orientation.x = q[0]; out.pose.pose.orientation.y = q[1]; out.pose.pose.orientation.z = q[2]; out.pose.pose.orientation.w = q[3];
            out.pose.pose.position.x = t[0]; out.pose.pose.position.y = t[1]; out.pose.pose.position.z = t[2];
            pubOdomAftMapped.publish(out);
            // trajectory line, LM:2284-2325
            Eigen::Matrix4f H = Eigen::Matrix4f::Identity();
            H.block<3, 3>(0, 0) = Eigen::Quaterniond(q[3], q[0], q[1], q[2]).toRotationMatrix().cast<float>();
            H(0, 3) = (float)t[0]; H(1, 3) = (float)t[1]; H(2, 3) = (float)t[2];
            if (init_flag) { H_init = H; init_flag = false; }
            H = H_init.inverse() * H;
            std::ofstream f(RESULT_PATH, std::ios::app);
            f.setf(std::ios::scientific, std::ios::floatfield);
            f.precision(6);
            for (int r = 0; r < 3; ++r) for (int c = 0; c < 4; ++c) f << H(r, c) << ((r == 2 && c == 3) ? "\n" : " ");
            geometry_msgs::PoseStamped ps;
            ps.header = out.header;
            ps.pose = out.pose.pose;
            path.header = out.header;
            path.poses.push_back(ps);
            pubPath.publish(path);
            static tf::TransformBroadcaster br;
            tf::Transform tr;
            tr.setOrigin(tf::Vector3(t[0], t[1], t[2]));
            tr.setRotation(tf::Quaternion(q[0], q[1], q[2], q[3]));
            br.sendTransform(tf::StampedTransform(tr, out.header.stamp, "rslidar", "/aft_mapped"));
        }
        std::this_thread::sleep_for(std::chrono::milliseconds(2));
    }
}

int main(int argc, char** argv)
{
    ros::init(argc, argv, "laserMapping");
    ros::NodeHandle nh;
    float lineRes = 0.4f, planeRes = 0.8f;
    int n_scans = 64;
    nh.param<float>("mapping_line_resolution", lineRes, 0.4);
    nh.param<float>("mapping_plane_resolution", planeRes, 0.8);
    nh.param<std::string>("RESULT_PATH", RESULT_PATH, " ");
    nh.param<int>("scan_line", n_scans, 64);
    ll_config cfg;
    ll_default_config(&cfg, n_scans);
    cfg.line_res = lineRes; cfg.plane_res = planeRes; cfg.enable_mapping = 1; cfg.max_points = 400000; cfg.map_capacity = 1 << 21;
    if (int rc = ll_create(&cfg, &g_ll)) { ROS_FATAL("lightloam_b200: %s", ll_strerror(rc)); return 1; }
    auto push = [](auto& q) { return [&q](const auto& m) { std::lock_guard<std::mutex> l(mBuf); q.push(m); }; };
    ros::Subscriber s1 = nh.subscribe<sensor_msgs::PointCloud2>("/laser_cloud_corner_last", 100, boost::function<void(const sensor_msgs::PointCloud2ConstPtr&)>(push(qCorner)));
    ros::Subscriber s2 = nh.subscribe<sensor_msgs::PointCloud2>("/laser_cloud_surf_last", 100, boost::function<void(const sensor_msgs::PointCloud2ConstPtr&)>(push(qSurf)));
    ros::Subscriber s3 = nh.subscribe<nav_msgs::Odometry>("/laser_odom_to_init", 100, boost::function<void(const nav_msgs::Odometry::ConstPtr&)>(push(qOdom)));
    ros::Subscriber s4 = nh.subscribe<sensor_msgs::PointCloud2>("/velodyne_cloud_3", 100, boost::function<void(const sensor_msgs::PointCloud2ConstPtr&)>(push(qFull)));
    pubOdomAftMapped = nh.advertise<nav_msgs::Odometry>("/aft_mapped_to_init", 100);
    pubPath = nh.advertise<nav_msgs::Path>("/aft_mapped_path", 100);
    std::thread worker{process};
    ros::spin();
    worker.join();
    ll_destroy(g_ll);
    return 0;
}
#endif  // LL_WITH_ROS
