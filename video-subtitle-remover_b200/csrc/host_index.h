// Host-side integer mask / index path (A1, A2, A6, T4 of SURVEY.md §8a) — plain C++, bit-exact against
// the reference functions; no CUDA involved.
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <numeric>
#include <tuple>
#include <vector>

namespace vsr {

// backend/tools/inpaint_tools.py:31-47 — cv2.rectangle(thickness=-1) fills the inclusive, clipped box.
inline void host_create_mask(uint8_t* mask, int H, int W, const int32_t* boxes, int n, int deviation) {
  std::memset(mask, 0, (size_t)H * W);
  for (int i = 0; i < n; ++i) {
    const int xmin = boxes[4 * i], xmax = boxes[4 * i + 1], ymin = boxes[4 * i + 2], ymax = boxes[4 * i + 3];
    int x1 = std::max(0, xmin - deviation), y1 = std::max(0, ymin - deviation);
    int x2 = xmax + deviation, y2 = ymax + deviation;
    int xa = std::min(x1, x2), xb = std::max(x1, x2), ya = std::min(y1, y2), yb = std::max(y1, y2);
    xa = std::max(xa, 0);
    ya = std::max(ya, 0);
    xb = std::min(xb, W - 1);
    yb = std::min(yb, H - 1);
    for (int y = ya; y <= yb; ++y)
      if (xa <= xb) std::memset(mask + (size_t)y * W + xa, 255, (size_t)(xb - xa + 1));
  }
}

struct Island {
  int top, bottom, center, area;
  long long order;  // first 2x2 block in block-raster order (OpenCV Spaghetti/BBDT label order)
};

// 8-connected components by run-based union-find.  Returns islands in cv2 label order.
inline std::vector<Island> host_components8(const uint8_t* mask, int H, int W) {
  struct Run { int y, x0, x1, parent; };
  std::vector<Run> runs;
  std::vector<int> row_start(H + 1, 0);
  for (int y = 0; y < H; ++y) {
    row_start[y] = (int)runs.size();
    const uint8_t* r = mask + (size_t)y * W;
    int x = 0;
    while (x < W) {
      if (r[x]) {
        int s = x;
        while (x < W && r[x]) ++x;
        runs.push_back({y, s, x - 1, (int)runs.size()});
      } else {
        ++x;
      }
    }
  }
  row_start[H] = (int)runs.size();
  auto find = [&](int a) {
    while (runs[a].parent != a) {
      runs[a].parent = runs[runs[a].parent].parent;
      a = runs[a].parent;
    }
    return a;
  };
  for (int y = 1; y < H; ++y) {
    int i = row_start[y - 1], iend = row_start[y];
    for (int j = row_start[y]; j < row_start[y + 1]; ++j) {
      // runs on the previous row overlapping [x0-1, x1+1] are 8-connected
      while (i < iend && runs[i].x1 < runs[j].x0 - 1) ++i;
      for (int k = i; k < iend && runs[k].x0 <= runs[j].x1 + 1; ++k) {
        int a = find(k), b = find(j);
        if (a != b) runs[std::max(a, b)].parent = std::min(a, b);
      }
    }
  }
  struct Acc { int top = 1 << 30, bot = -1, area = 0; double sumy = 0; long long order = (1LL << 62); };
  std::vector<Acc> acc(runs.size());
  const long long bw = W / 2 + 2;
  for (size_t i = 0; i < runs.size(); ++i) {
    Acc& a = acc[find((int)i)];
    const Run& r = runs[i];
    const int len = r.x1 - r.x0 + 1;
    a.top = std::min(a.top, r.y);
    a.bot = std::max(a.bot, r.y);
    a.area += len;
    a.sumy += (double)r.y * len;
    a.order = std::min(a.order, (long long)(r.y / 2) * bw + r.x0 / 2);
  }
  std::vector<Island> out;
  for (size_t i = 0; i < runs.size(); ++i) {
    if (runs[i].parent != (int)i) continue;
    const Acc& a = acc[i];
    out.push_back({a.top, a.bot + 1, (int)(a.sumy / a.area), a.area, a.order});
  }
  std::sort(out.begin(), out.end(), [](const Island& a, const Island& b) { return a.order < b.order; });
  return out;
}

// backend/tools/inpaint_tools.py:49-242.  Returns (ymin, ymax, xmin, xmax) strips.
inline std::vector<std::array<int, 4>> host_inpaint_areas(int W, int H, int h, const uint8_t* mask, int multiple) {
  std::vector<std::array<int, 4>> areas;
  std::vector<Island> isl;
  for (const Island& i : host_components8(mask, H, W))
    if (i.area >= 10) isl.push_back(i);  // :88
  if (isl.empty()) return areas;
  std::stable_sort(isl.begin(), isl.end(), [](const Island& a, const Island& b) { return a.center < b.center; });  // :99
  std::vector<uint8_t> row_any(H, 0);
  for (int y = 0; y < H; ++y) {
    const uint8_t* r = mask + (size_t)y * W;
    for (int x = 0; x < W; ++x)
      if (r[x]) { row_any[y] = 1; break; }
  }
  std::vector<std::vector<Island>> groups(1, std::vector<Island>{isl[0]});
  for (size_t i = 1; i < isl.size(); ++i) {
    auto& cur = groups.back();
    int lo = 1 << 30, hi = -1;
    for (auto& g : cur) { lo = std::min(lo, g.top); hi = std::max(hi, g.bottom); }
    const int nlo = std::min(lo, isl[i].top), nhi = std::max(hi, isl[i].bottom);
    bool connected = true;
    if (hi < isl[i].top) {  // :119-125
      connected = false;
      for (int y = hi; y < isl[i].top; ++y)
        if (row_any[y]) { connected = true; break; }
    }
    if (nhi - nlo <= h && connected) cur.push_back(isl[i]);
    else groups.push_back(std::vector<Island>{isl[i]});
  }
  auto floordiv = [](long long a, long long b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); };
  for (auto& g : groups) {
    int lo = 1 << 30, hi = -1;
    long long csum = 0;
    for (auto& i : g) { lo = std::min(lo, i.top); hi = std::max(hi, i.bottom); csum += i.center; }
    const int cy = (int)floordiv(csum, (long long)g.size());
    const int half = h / 2;
    auto place = [&](int c, int& ymin, int& ymax) {
      ymin = std::max(0, c);
      ymax = ymin + h;
      if (ymax > H) { ymax = H; ymin = std::max(0, H - h); }
    };
    int ymin, ymax;
    place(cy - half, ymin, ymax);
    if (ymin > lo || ymax < hi) {  // :164-184
      if (hi - lo <= h) place(lo, ymin, ymax);
      else place((lo + hi) / 2 - half, ymin, ymax);
    }
    int xmin = 0, xmax = W;
    if (multiple > 1) {  // :189-235
      const int height = ymax - ymin;
      const int rem = height % multiple;
      if (rem != 0) {
        const int adj = multiple - rem;
        const double c = (ymin + ymax) / 2.0;
        if (ymin - adj / 2.0 >= 0 && ymax + adj / 2.0 <= H) {
          const int a = (int)(c - height / 2.0 - adj / 2.0), b = (int)(c + height / 2.0 + adj / 2.0);
          ymin = a; ymax = b;
        } else if (height > multiple) {
          const int a = (int)(c - (height - rem) / 2.0), b = (int)(c + (height - rem) / 2.0);
          ymin = a; ymax = b;
        } else {
          if (ymax + adj <= H) ymax += adj;
          else if (ymin - adj >= 0) ymin -= adj;
          else if (height > multiple) ymax = ymin + height - rem;
        }
      }
      const int width = xmax - xmin;
      const int remw = width % multiple;
      if (remw != 0) {
        const double cx = (xmin + xmax) / 2.0;
        const int a = (int)(cx - (width - remw) / 2.0), b = (int)(cx + (width - remw) / 2.0);
        xmin = a; xmax = b;
      }
    }
    std::array<int, 4> area{ymin, ymax, xmin, xmax};
    if (std::find(areas.begin(), areas.end(), area) == areas.end()) areas.push_back(area);
  }
  return areas;
}

// backend/tools/inpaint_tools.py:7-29
inline std::vector<int> host_batch_sizes(int n, int max_bs) {
  std::vector<int> out;
  if (n <= 0 || max_bs <= 0) return out;
  int bs = max_bs, nb = n / bs;
  while ((double)(n % bs) < bs / 2.0 && bs > 1) {
    --bs;
    nb = n / bs;
  }
  for (int i = 0; i < nb; ++i) out.push_back(bs);
  if (nb * bs < n) out.push_back(n - nb * bs);
  return out;
}

struct Window {
  std::vector<int> neighbors, refs;
};
// backend/inpaint/sttn_auto_inpaint.py:142-146 and get_ref_index :107-120
inline std::vector<Window> host_window_schedule(int T, int stride, int ref_length) {
  std::vector<Window> out;
  for (int f = 0; f < T; f += stride) {
    Window w;
    const int lo = std::max(0, f - stride), hi = std::min(T, f + stride + 1);
    for (int i = lo; i < hi; ++i) w.neighbors.push_back(i);
    for (int i = 0; i < T; i += ref_length)
      if (i < lo || i >= hi) w.refs.push_back(i);
    out.push_back(std::move(w));
  }
  return out;
}

// cv::resize INTER_LINEAR tap tables (imgproc/resize.cpp): scale = 1/(dst/src) in double,
// f = float((d+0.5)*scale-0.5); horizontal taps clamp the coefficient at the borders, vertical taps keep
// the fraction and clamp the row indices.  Weights: saturate_cast<short>(w*2048) (round-half-even).
struct HostTaps {
  std::vector<int> i0, i1;
  std::vector<float> a;
  std::vector<short> w0, w1;
};
inline HostTaps host_resize_taps(int src_n, int dst_n, bool vertical) {
  HostTaps t;
  t.i0.resize(dst_n); t.i1.resize(dst_n); t.a.resize(dst_n); t.w0.resize(dst_n); t.w1.resize(dst_n);
  const double scale = 1.0 / ((double)dst_n / (double)src_n);
  for (int d = 0; d < dst_n; ++d) {
    float f = (float)((d + 0.5) * scale - 0.5);
    int s = (int)std::floor(f);
    float a = f - (float)s;
    int i0, i1;
    if (vertical) {
      i0 = std::min(std::max(s, 0), src_n - 1);
      i1 = std::min(std::max(s + 1, 0), src_n - 1);
    } else {
      if (s < 0) { s = 0; a = 0.f; }
      if (s >= src_n - 1) { s = src_n - 1; a = 0.f; }
      i0 = s;
      i1 = std::min(s + 1, src_n - 1);
    }
    t.i0[d] = i0; t.i1[d] = i1; t.a[d] = a;
    t.w0[d] = (short)std::nearbyintf((1.0f - a) * 2048.0f);
    t.w1[d] = (short)std::nearbyintf(a * 2048.0f);
  }
  return t;
}

}  // namespace vsr
