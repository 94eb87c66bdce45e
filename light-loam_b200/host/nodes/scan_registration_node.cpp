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
static std::vector<float> bFull, bSharp, bLessSharp, bFlat, bLessFlat;

static void publish(ros::Publisher& pub, const ll_cloud_out& c, const std_msgs::Header& h)
{
    pcl::PointCloud<pcl::PointXYZI> cloud;
    cloud.resize(c.n);
    for (int i = 0; i < c.n; ++i) {
        cloud[i].x = c.xyzi[4 * i]; cloud[i].y = c.xyzi[4 * i + 1]; cloud[i].z = c.xyzi[4 * i + 2]; cloud[i].intensity = c.xyzi[4 * i + 3];
    }
    sensor_msgs::PointCloud2 msg;
    pcl::toROSMsg(cloud, msg);
    msg.header.stamp = h.stamp;
    msg.header.frame_id = h.frame_id;
    pub.publish(msg);
}

static void laserCloudHandler(const sensor_msgs::PointCloud2ConstPtr& in)
{
    pcl::PointCloud<pcl::PointXYZ> cloud;
    pcl::fromROSMsg(*in, cloud);  // SR:105-106; NaN / range filters run inside the call
    ll_cloud_view scan{reinterpret_cast<const float*>(cloud.points.data()), (int)cloud.size(), (int)sizeof(pcl::PointXYZ)};
    ll_cloud_out full{bFull.data(), 0, (int)bFull.size() / 4}, sharp{bSharp.data(), 0, (int)bSharp.size() / 4},
        lsharp{bLessSharp.data(), 0, (int)bLessSharp.size() / 4}, flat{bFlat.data(), 0, (int)bFlat.size() / 4},
        lflat{bLessFlat.data(), 0, (int)bLessFlat.size() / 4};
    const int rc = ll_extract_features(g_ll, scan, &full, &sharp, &lsharp, &flat, &lflat, nullptr, nullptr, nullptr, nullptr, nullptr);
    if (rc < 0) { ROS_WARN("lightloam_b200: %s", ll_strerror(rc)); return; }
    publish(pubFull, full, in->header);
    publish(pubSharp, sharp, in->header);
    publish(pubLessSharp, lsharp, in->header);
    publish(pubFlat, flat, in->header);
    publish(pubLessFlat, lflat, in->header);
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
    bFull.resize(4 * 400000); bLessFlat.resize(4 * 400000); bSharp.resize(4 * n_scans * 12); bLessSharp.resize(4 * n_scans * 120); bFlat.resize(4 * n_scans * 24);
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
