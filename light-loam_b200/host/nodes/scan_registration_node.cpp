// ascanRegistration with the hot path replaced by ll_extract_features (scanRegistration.cpp:100-377).
// Same subscription (/rslidar_points, SR:453), same five publications with the input stamp / frame (SR:382-410),
// same params (scan_line, minimum_range, lowerBound, upBound; SR:435-443).
#ifdef LL_WITH_ROS
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#include <pcl_conversions/pcl_conversions.h>
#include <ros/ros.h>
#include <sensor_msgs/PointCloud2.h>

#include "lightloam_b200.h"

static ll_ctx* g_ll = nullptr;
static ros::Publisher pubFull, pubSharp, pubLessSharp, pubFlat, pubLessFlat;
static sensor_msgs::PointCloud2 g_template;   // fields / point_step of pcl::toROSMsg(PointCloud<PointXYZI>), built once

// One of the five clouds of the last extraction, packed on the device into the PointXYZI wire layout
// (ll_fetch_pointcloud2) straight into msg.data: no per-point work on the host (SR:382-410).
static void publish(ros::Publisher& pub, int which, int cap, const std_msgs::Header& h)
{
    sensor_msgs::PointCloud2 msg = g_template;
    msg.data.resize((size_t)cap * 32);
    int n = 0;
    if (ll_fetch_pointcloud2(g_ll, which, msg.data.data(), cap, &n) < 0) return;
    msg.data.resize((size_t)n * 32);
    msg.width = n; msg.height = 1; msg.row_step = n * 32; msg.is_dense = true;
    msg.header.stamp = h.stamp;
    msg.header.frame_id = h.frame_id;
    pub.publish(msg);
}

static void laserCloudHandler(const sensor_msgs::PointCloud2ConstPtr& in)
{
    // SR:105-106 (fromROSMsg) + SR:109-110: the payload is read in place - x, y, z are the first three fp32 fields of
    // the sensor drivers' layouts; stride = point_step; the NaN / range filters run inside the call
    ll_cloud_view scan{reinterpret_cast<const float*>(in->data.data()), (int)(in->width * in->height), (int)in->point_step};
    const int rc = ll_extract_features(g_ll, scan, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
    if (rc < 0) { ROS_WARN("lightloam_b200: %s", ll_strerror(rc)); return; }
    ll_stats st;
    ll_get_last_stats(g_ll, &st);
    publish(pubFull, 0, st.n_full, in->header);
    publish(pubSharp, 1, st.n_sharp, in->header);
    publish(pubLessSharp, 2, st.n_less_sharp, in->header);
    publish(pubFlat, 3, st.n_flat, in->header);
    publish(pubLessFlat, 4, st.n_less_flat, in->header);
}

int main(int argc, char** argv)
{
    ros::init(argc, argv, "scanRegistration");
    ros::NodeHandle nh;
    int n_scans = 16;
    double min_range = 0.1;
    float lower = -24.9f, upper = 2.f;
    nh.param<int>("scan_line", n_scans, 16);
    nh.param<double>("minimum_range", min_range, 0.1);
    nh.param<float>("lowerBound", lower, -24.9);
    nh.param<float>("upBound", upper, 2);
    if (n_scans != 16 && n_scans != 32 && n_scans != 64) return 0;  // SR:447-451
    ll_config cfg;
    ll_default_config(&cfg, n_scans);
    cfg.minimum_range = (float)min_range; cfg.lower_bound = lower; cfg.up_bound = upper; cfg.max_points = 400000;  // SR:34
    if (int rc = ll_create(&cfg, &g_ll)) { ROS_FATAL("lightloam_b200: %s", ll_strerror(rc)); return 1; }
    { pcl::PointCloud<pcl::PointXYZI> empty; pcl::toROSMsg(empty, g_template); }
    ros::Subscriber sub = nh.subscribe<sensor_msgs::PointCloud2>("/rslidar_points", 100, laserCloudHandler);
    pubFull = nh.advertise<sensor_msgs::PointCloud2>("/velodyne_cloud_2", 100);
    pubSharp = nh.advertise<sensor_msgs::PointCloud2>("/laser_cloud_sharp", 100);
    pubLessSharp = nh.advertise<sensor_msgs::PointCloud2>("/laser_cloud_less_sharp", 100);
    pubFlat = nh.advertise<sensor_msgs::PointCloud2>("/laser_cloud_flat", 100);
    pubLessFlat = nh.advertise<sensor_msgs::PointCloud2>("/laser_cloud_less_flat", 100);
    ros::spin();
    ll_destroy(g_ll);
    return 0;
}
#endif  // LL_WITH_ROS
