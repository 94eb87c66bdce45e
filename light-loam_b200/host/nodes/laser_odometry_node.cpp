// alaserOdometry with the loop body replaced by ll_odometry_step (laserOdometry.cpp:425-896).
// Same five subscriptions (LO:354-362), same publications: /laser_odom_to_init (frame "rslidar", child "/laser_odom",
// LO:837-850), /laser_odom_path, and every skipFrameNum frames corner_last / surf_last / velodyne_cloud_3 (LO:898-919).
#ifdef LL_WITH_ROS
#include <geometry_msgs/PoseStamped.h>
#include <nav_msgs/Odometry.h>
#include <nav_msgs/Path.h>
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#include <pcl_conversions/pcl_conversions.h>
#include <ros/ros.h>
#include <sensor_msgs/PointCloud2.h>

#include <mutex>
#include <queue>

#include "lightloam_b200.h"

static std::queue<sensor_msgs::PointCloud2ConstPtr> qSharp, qLessSharp, qFlat, qLessFlat, qFull;
static std::mutex mBuf;
#define HANDLER(name, q) static void name(const sensor_msgs::PointCloud2ConstPtr& m) { std::lock_guard<std::mutex> l(mBuf); q.push(m); }
HANDLER(hSharp, qSharp) HANDLER(hLessSharp, qLessSharp) HANDLER(hFlat, qFlat) HANDLER(hLessFlat, qLessFlat) HANDLER(hFull, qFull)


int main(int argc, char** argv)
{
    ros::init(argc, argv, "laserOdometry");
    ros::NodeHandle nh;
    int n_scans = 16, skipFrameNum = 2;
    nh.param<int>("scan_line", n_scans, 16);
    nh.param<int>("mapping_skip_frame", skipFrameNum, 2);
    ll_config cfg;
    ll_default_config(&cfg, n_scans);
    cfg.max_points = 400000;
    ll_ctx* ll = nullptr;
    if (int rc = ll_create(&cfg, &ll)) { ROS_FATAL("lightloam_b200: %s", ll_strerror(rc)); return 1; }
    ros::Subscriber s1 = nh.subscribe<sensor_msgs::PointCloud2>("/laser_cloud_sharp", 100, hSharp);
    ros::Subscriber s2 = nh.subscribe<sensor_msgs::PointCloud2>("/laser_cloud_less_sharp", 100, hLessSharp);
    ros::Subscriber s3 = nh.subscribe<sensor_msgs::PointCloud2>("/laser_cloud_flat", 100, hFlat);
    ros::Subscriber s4 = nh.subscribe<sensor_msgs::PointCloud2>("/laser_cloud_less_flat", 100, hLessFlat);
    ros::Subscriber s5 = nh.subscribe<sensor_msgs::PointCloud2>("/velodyne_cloud_2", 100, hFull);
    ros::Publisher pubCornerLast = nh.advertise<sensor_msgs::PointCloud2>("/laser_cloud_corner_last", 100);
    ros::Publisher pubSurfLast = nh.advertise<sensor_msgs::PointCloud2>("/laser_cloud_surf_last", 100);
    ros::Publisher pubFull = nh.advertise<sensor_msgs::PointCloud2>("/velodyne_cloud_3", 100);
    ros::Publisher pubOdom = nh.advertise<nav_msgs::Odometry>("/laser_odom_to_init", 100);
    ros::Publisher pubPath = nh.advertise<nav_msgs::Path>("/laser_odom_path", 100);
    nav_msgs::Path path;
    int frameCount = 0;
    ros::Rate rate(100);
    while (ros::ok()) {
        ros::spinOnce();
        if (!qSharp.empty() && !qLessSharp.empty() && !qFlat.empty() && !qLessFlat.empty() && !qFull.empty()) {
            mBuf.lock();
            const ros::Time stamp = qLessFlat.front()->header.stamp;
            if (qSharp.front()->header.stamp != qFull.front()->header.stamp || qLessSharp.front()->header.stamp != qFull.front()->header.stamp ||
                qFlat.front()->header.stamp != qFull.front()->header.stamp || stamp != qFull.front()->header.stamp) {
                mBuf.unlock();
                ROS_BREAK();  // LO:394-401
            }
            sensor_msgs::PointCloud2ConstPtr mLessSharp = qLessSharp.front(), mLessFlat = qLessFlat.front(), mFull = qFull.front();
            sensor_msgs::PointCloud2ConstPtr mSharp = qSharp.front(), mFlat = qFlat.front();
            qSharp.pop(); qLessSharp.pop(); qFlat.pop(); qLessFlat.pop(); qFull.pop();
            mBuf.unlock();
            double q[4], t[3], ql[4], tl[3];
            // the PointXYZI payloads are read in place: point_step 32, x,y,z at 0, intensity at byte 16 (LO:403-423 fromROSMsg)
            auto view = [](const sensor_msgs::PointCloud2& m) {
                return ll_cloud_view{reinterpret_cast<const float*>(m.data.data()), (int)(m.width * m.height), (int)m.point_step};
            };
            const ll_cloud_view va = view(*mSharp), vb = view(*mLessSharp), vc = view(*mFlat), vd = view(*mLessFlat);
            const int rc = ll_odometry_step(ll, va, vb, vc, vd, q, t, ql, tl);
            if (rc < 0) { ROS_WARN("lightloam_b200: %s", ll_strerror(rc)); continue; }
            nav_msgs::Odometry odom;
            odom.header.frame_id = "rslidar";
            odom.child_frame_id = "/laser_odom";
            odom.header.stamp = stamp;
            odom.pose.pose.orientation.x = q[0]; odom.pose.pose.orientation.y = q[1]; odom.pose.pose.orientation.z = q[2]; odom.pose.pose.orientation.w = q[3];
            odom.pose.pose.position.x = t[0]; odom.pose.pose.position.y = t[1]; odom.pose.pose.position.z = t[2];
            pubOdom.publish(odom);
            geometry_msgs::PoseStamped ps;
            ps.header = odom.header;
            ps.pose = odom.pose.pose;
            path.header.stamp = odom.header.stamp;
            path.header.frame_id = "rslidar";
            path.poses.push_back(ps);
            pubPath.publish(path);
            if (frameCount % skipFrameNum == 0) {  // LO:898-919: the clouds just swapped in are this frame's less-sharp / less-flat
                frameCount = 0;
                sensor_msgs::PointCloud2 m1 = *mLessSharp, m2 = *mLessFlat, m3 = *mFull;
                m1.header.stamp = m2.header.stamp = m3.header.stamp = stamp;
                m1.header.frame_id = m2.header.frame_id = m3.header.frame_id = "/camera";
                pubCornerLast.publish(m1);
                pubSurfLast.publish(m2);
                pubFull.publish(m3);
            }
            frameCount++;
        }
        rate.sleep();
    }
    ll_destroy(ll);
    return 0;
}
#endif  // LL_WITH_ROS
