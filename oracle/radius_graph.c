/* TEST INFRASTRUCTURE ONLY -- C restatement of PointVS generate_edges
 * (/root/reference/point_vs/preprocessing/preprocessing.py:68-155, no prune).
 *
 * Distances follow scipy's euclidean cdist (preprocessing.py:108) exactly:
 * sqrt((dx*dx + dy*dy) + dz*dz) in IEEE double with NO fused multiply-add;
 * build with -ffp-contract=off (oracle/Makefile does).  The output order is
 * the reference's: the inter list (d < inter_radius, d > 1e-7, bp differs),
 * then the intra list (d < intra_radius, d > 1e-7, any pair), each row-major.
 *
 * Two-call protocol: call with cap == 0 to obtain the edge count, then with
 * buffers of that size.  Returns the number of edges (or -1 on overflow).
 */
#include <math.h>
#include <stddef.h>

static double dist(const double *c, long i, long j)
{
    double dx = c[3 * i] - c[3 * j];
    double dy = c[3 * i + 1] - c[3 * j + 1];
    double dz = c[3 * i + 2] - c[3 * j + 2];
    return sqrt((dx * dx + dy * dy) + dz * dz);
}

long pvs_oracle_radius_graph(const double *coords, const int *bp, long n,
                             double inter_radius, double intra_radius,
                             long *row, long *col, int *attr, long cap)
{
    long e = 0;
    for (int pass = 0; pass < 2; ++pass) {
        double radius = pass == 0 ? inter_radius : intra_radius;
        for (long i = 0; i < n; ++i) {
            for (long j = 0; j < n; ++j) {
                double d = dist(coords, i, j);
                if (!(d < radius && d > 1e-7))
                    continue;
                int a;
                if (pass == 0) {
                    if (bp[i] == bp[j])
                        continue;
                    a = ((bp[i] == 0 && bp[j] == 1) ||
                         (bp[i] == 1 && bp[j] == 0)) ? 1 : 0;
                } else {
                    a = (bp[i] == 1 && bp[j] == 1) ? 2 : 0;
                }
                if (cap > 0) {
                    if (e >= cap)
                        return -1;
                    row[e] = i;
                    col[e] = j;
                    attr[e] = a;
                }
                ++e;
            }
        }
    }
    return e;
}
