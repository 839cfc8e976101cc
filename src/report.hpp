// QC report of the tgsfilter host: the data shaping of the reference (Get_qual_Dis,
// Get_plot_line_data, Get_length_Dis, Get_N50, limitDecimalPlaces: T.cpp:2584-2921) and the two
// machine-readable parts of its HTML report — the `var data = {...}` block (genPlotData,
// include/report.cpp:597-628) and the summary table (genTable, include/report.cpp:631-668) —
// written from the counter block libtgsf_cuda returns.  Numbers are produced with the reference's
// own float expressions and stream formatting so the data block and the table compare equal to
// the reference's; the page around them (styles, chart script) is this repo's own minimal one —
// the reference embeds a 1 MB third-party chart library that is not vendored here, the page loads
// it from a CDN when opened online and degrades to the table + raw data otherwise.
//
// Quirks reproduced on purpose: bins whose start is not a plotted x are merged into
// x[0] through default-constructed map entries (T.cpp:2681-2683); the 3' tables are added twice into
// the merged counts, halving the plotted 3' qualities (T.cpp:2718-2725).
#pragma once
#include <cmath>
#include <cstdint>
#include <ctime>
#include <fstream>
#include <sstream>
#include <string>
#include <unordered_map>
#include <vector>

#include "../include/tgsf.h"

namespace report {

struct LineY { std::string bases; std::vector<float> data; };
struct LinePlot { std::vector<int> x; std::vector<LineY> y; };
struct LenDis { std::vector<int> x; std::vector<uint64_t> y; };
struct QualDis { std::vector<int> x; std::vector<float> y; };

struct Side { // one column of the report: raw or clean / downsampled
    LenDis lenDis;
    QualDis qualDis;
    LinePlot readsQual, readsQual5p, readsQual3p, basesContents, basesContents5p, basesContents3p;
    std::string tab[9]; // reads, bases, GC, min, max, mean, median, N50, mean quality
};

// [rows][5] table view into the counter block
struct Table {
    const uint64_t *p;
    uint32_t rows;
    uint64_t at(uint32_t r, int c) const { return r < rows ? p[(uint64_t)r * 5 + c] : 0; }
};

inline std::string limitDecimalPlaces(double num, int places) { // T.cpp:2911-2921
    std::string str = std::to_string(num);
    size_t dot = str.find('.');
    if (dot != std::string::npos && str.size() - dot > (size_t)(places + 1)) str = str.substr(0, dot + places + 1);
    return str;
}

inline int n50(const std::vector<int> &lens, uint64_t bases) { // Get_N50, T.cpp:2900-2909 (lens sorted)
    uint64_t acc = 0;
    for (int i = (int)lens.size() - 1; i >= 0; --i) {
        acc += (uint64_t)lens[(size_t)i];
        if (acc >= bases / 2) return lens[(size_t)i];
    }
    return 0;
}

inline void length_dis(const std::vector<int> &lens, LenDis &out) { // Get_length_Dis, T.cpp:2852-2898
    int i = lens.front();
    const int maxLen = lens.back();
    std::vector<uint64_t> indexs;
    while (i < maxLen) {
        indexs.push_back((uint64_t)i);
        int step = i / 20;
        if (step < 100) step = 100;
        i += step;
    }
    indexs.push_back((uint64_t)maxLen);
    const int vecMax = (int)indexs.size() - 1;
    std::unordered_map<int, int> seqIndex, seqNum;
    for (int a = 0; a < vecMax; ++a)
        for (int x = (int)indexs[(size_t)a]; x < (int)indexs[(size_t)a + 1]; ++x) seqIndex[x] = (int)indexs[(size_t)a];
    if (vecMax >= 1) seqIndex[maxLen] = (int)indexs[(size_t)vecMax - 1];
    for (int len : lens) seqNum[seqIndex[len]]++;
    out.x.resize((size_t)vecMax);
    out.y.resize((size_t)vecMax);
    for (int a = 0; a < vecMax; ++a) {
        out.x[(size_t)a] = (int)indexs[(size_t)a];
        out.y[(size_t)a] = (uint64_t)seqNum[(int)indexs[(size_t)a]];
    }
}

inline void qual_dis(const uint64_t *hist, uint64_t baseNum, QualDis &out) { // Get_qual_Dis, T.cpp:2584-2605
    int maxQual = 0;
    for (int j = 0; j < TGSF_QUAL_HIST_N; ++j)
        if (hist[j] > 0) maxQual = j;
    out.x.resize((size_t)maxQual + 1);
    out.y.resize((size_t)maxQual + 1);
    for (int i = 0; i <= maxQual; ++i) {
        out.x[(size_t)i] = i;
        out.y[(size_t)i] = static_cast<float>(hist[i] * 100) / baseNum;
    }
}

// Get_plot_line_data, T.cpp:2607-2850.  binCnt/binQual: per-100 bp tables, c5/q5/c3/q3: end tables.
inline void plot_line_data(int endLen, int maxLen, uint64_t bases, const Table &binCnt, const Table &binQual,
                           const Table &c5, const Table &q5, const Table &c3, const Table &q3, Side &S, float &gc,
                           float &meanQual) {
    const char *names[5] = {"A", "T", "G", "C", "Mean"};
    const int baseNum = 5;
    std::vector<uint64_t> indexs;
    for (uint64_t i = 1; i < (uint64_t)maxLen;) {
        indexs.push_back(i);
        uint64_t step = i / 20;
        if (step < 100) step = 100;
        i += step;
    }
    indexs.push_back((uint64_t)maxLen);
    const uint64_t vecMax = indexs.size() - 1;
    std::unordered_map<uint64_t, uint64_t> vecIndexs, lenIndexs;
    for (uint64_t i = 0; i < vecMax; ++i) {
        vecIndexs[indexs[i]] = i;
        for (uint64_t x = indexs[i]; x < indexs[i + 1]; ++x) lenIndexs[x] = indexs[i];
    }
    lenIndexs[(uint64_t)maxLen] = indexs.back();

    std::vector<std::vector<uint64_t>> cntM(vecMax, std::vector<uint64_t>(baseNum)), qualM(vecMax, std::vector<uint64_t>(baseNum));
    std::vector<std::vector<uint64_t>> c5M((size_t)endLen, std::vector<uint64_t>(baseNum)), q5M = c5M, c3M = c5M, q3M = c5M;
    // the reference's vectors hold int(maxLen/100)+1 bins (T.cpp:1445-1449)
    const uint64_t nbins = (uint64_t)maxLen / 100 + 1;
    for (uint64_t j = 0; j < nbins && vecMax > 0; ++j) {
        const uint64_t vi = vecIndexs[lenIndexs[j * 100 + 1]]; // missing keys read as 0, like operator[]
        for (int x = 0; x < baseNum; ++x) {
            cntM[vi][(size_t)x] += binCnt.at((uint32_t)j, x);
            qualM[vi][(size_t)x] += binQual.at((uint32_t)j, x);
        }
    }
    for (int j = 0; j < endLen; ++j)
        for (int x = 0; x < baseNum; ++x) {
            c5M[(size_t)j][(size_t)x] += c5.at((uint32_t)j, x);
            q5M[(size_t)j][(size_t)x] += q5.at((uint32_t)j, x);
            c3M[(size_t)j][(size_t)x] += c3.at((uint32_t)j, x);
            q3M[(size_t)j][(size_t)x] += q3.at((uint32_t)j, x);
            // T.cpp:2718-2725: the 3' counts are added a second time inside the quality loop, but only
            // for rows the quality table has (it has none for FASTA input)
            if (q3.rows > (uint32_t)j) c3M[(size_t)j][(size_t)x] += c3.at((uint32_t)j, x);
        }

    auto init = [&](LinePlot &P, size_t n, int series) {
        P.x.resize(n);
        P.y.resize((size_t)series);
        for (int i = 0; i < series; ++i) { P.y[(size_t)i].bases = names[i]; P.y[(size_t)i].data.assign(n, 0.0f); }
    };
    auto fill = [&](LinePlot &Q, LinePlot &Cn, const std::vector<std::vector<uint64_t>> &cnt,
                    const std::vector<std::vector<uint64_t>> &qual, size_t i) {
        for (int j = 0; j < baseNum - 1; ++j) {
            Q.y[(size_t)j].data[i] = cnt[i][(size_t)j] > 0 ? static_cast<float>(qual[i][(size_t)j]) / cnt[i][(size_t)j] : 0;
            Cn.y[(size_t)j].data[i] = cnt[i][baseNum - 1] > 0 ? static_cast<float>(cnt[i][(size_t)j] * 100) / cnt[i][baseNum - 1] : 0;
        }
        Q.y[baseNum - 1].data[i] = cnt[i][baseNum - 1] > 0 ? static_cast<float>(qual[i][baseNum - 1]) / cnt[i][baseNum - 1] : 0;
    };
    init(S.readsQual, vecMax, baseNum);
    init(S.basesContents, vecMax, baseNum - 1);
    uint64_t gcSum = 0, qualSum = 0;
    for (uint64_t i = 0; i < vecMax; ++i) {
        S.readsQual.x[i] = (int)indexs[i];
        S.basesContents.x[i] = (int)indexs[i];
        gcSum += cntM[i][2] + cntM[i][3];
        qualSum += qualM[i][baseNum - 1];
        fill(S.readsQual, S.basesContents, cntM, qualM, i);
    }
    gc = static_cast<float>(gcSum * 100) / bases;
    meanQual = static_cast<float>(qualSum) / bases;
    init(S.readsQual5p, (size_t)endLen, baseNum);
    init(S.basesContents5p, (size_t)endLen, baseNum - 1);
    init(S.readsQual3p, (size_t)endLen, baseNum);
    init(S.basesContents3p, (size_t)endLen, baseNum - 1);
    for (int i = 0; i < endLen; ++i) {
        S.readsQual5p.x[(size_t)i] = S.basesContents5p.x[(size_t)i] = i + 1;
        S.readsQual3p.x[(size_t)i] = S.basesContents3p.x[(size_t)i] = i + 1;
        fill(S.readsQual5p, S.basesContents5p, c5M, q5M, (size_t)i);
        fill(S.readsQual3p, S.basesContents3p, c3M, q3M, (size_t)i);
    }
}

// One report column from (sorted copy of) the lengths + the matching tables (T.cpp:3146-3173).
inline void build_side(std::vector<int> lens, uint64_t bases, int endLen, bool has_qual, const uint64_t *C,
                       const tgsf_counter_layout &L, bool clean, int qual_places, Side &S) {
    const size_t num = lens.size();
    if (num == 0) return;
    std::sort(lens.begin(), lens.end());
    const uint32_t bins = L.max_bins, bc = L.bc_len;
    const uint32_t qrows = has_qual ? bins : 0, qrows_end = has_qual ? bc : 0;
    Table binCnt{C + (clean ? L.clean_bin_cnt : L.raw_bin_cnt), bins}, binQual{C + (clean ? L.clean_bin_qual : L.raw_bin_qual), qrows};
    Table c5{C + (clean ? L.clean5p_cnt : L.raw5p_cnt), bc}, q5{C + (clean ? L.clean5p_qual : L.raw5p_qual), qrows_end};
    Table c3{C + (clean ? L.clean3p_cnt : L.raw3p_cnt), bc}, q3{C + (clean ? L.clean3p_qual : L.raw3p_qual), qrows_end};
    float gc = 0, mq = 0;
    plot_line_data(endLen, lens.back(), bases, binCnt, binQual, c5, q5, c3, q3, S, gc, mq);
    S.tab[0] = std::to_string((int)num);
    S.tab[1] = std::to_string(bases);
    S.tab[2] = limitDecimalPlaces(std::round(gc * 1000) / 1000.0, 3);
    S.tab[3] = std::to_string(lens.front());
    S.tab[4] = std::to_string(lens.back());
    S.tab[5] = std::to_string((int)(bases / num));
    S.tab[6] = std::to_string(lens[num / 2]);
    S.tab[7] = std::to_string(n50(lens, bases));
    S.tab[8] = limitDecimalPlaces(std::round(mq * 1000) / 1000.0, qual_places);
    length_dis(lens, S.lenDis);
    qual_dis(C + (clean ? L.clean_hist : L.raw_hist), bases, S.qualDis);
}

template <typename T>
inline std::string json_array(const std::vector<T> &v) { // vectorNumToJson, include/report.cpp:18-33
    std::stringstream ss;
    ss << "[";
    for (size_t i = 0; i < v.size(); ++i) {
        ss << v[i];
        if (i + 1 < v.size()) ss << ",";
    }
    ss << "]";
    return ss.str();
}

inline int title_gap(int maxY) { // getYTitleGap, include/report.cpp:547-557
    std::string s = std::to_string(maxY);
    int size = (int)s.length();
    for (size_t i = 3; i < s.length(); i += 3) size += 1;
    return size * 5 + 30;
}
template <typename V> inline int max_y(const std::vector<V> &y) { // getMaxY: int maxY; if (value > maxY) maxY = value
    int m = 0;
    for (auto &v : y)
        if (v > (V)m) m = (int)v;
    return m;
}
inline int max_y(const LinePlot &p) {
    int m = 0;
    for (auto &item : p.y)
        for (float v : item.data)
            if (v > m) m = (int)v;
    return m;
}

inline void put(std::ostream &o, const char *key, const LenDis &v) {
    o << "" << key << ": {" << "x: " << json_array(v.x) << ",\n" << "y: " << json_array(v.y) << ",\n"
      << "yTitleGap: " << title_gap(max_y(v.y)) << ",\n" << "},\n";
}
inline void put(std::ostream &o, const char *key, const QualDis &v) {
    o << "" << key << ": {" << "x: " << json_array(v.x) << ",\n" << "y: " << json_array(v.y) << ",\n"
      << "yTitleGap: " << title_gap(max_y(v.y)) << ",\n" << "},\n";
}
inline void put(std::ostream &o, const char *key, const LinePlot &v) {
    o << "" << key << ": {" << "x: " << json_array(v.x) << "," << "y: [";
    for (auto &item : v.y) o << "{" << "name:\"" << item.bases << "\", " << "data:" << json_array(item.data) << ", " << "}, ";
    o << "], " << "yTitleGap:" << title_gap(max_y(v)) << ", " << "},";
}

// qcType: "00","02","01" (fasta) / "10","12","11" (fastq): [1] 0 = QC only, 1 = downsample only, 2 = filter
inline void write_html(const std::string &path, const std::string &qcType, const Side &raw, const Side &clean) {
    std::ofstream o(path);
    const bool fq = qcType[0] == '1', show_raw = qcType[1] != '1', show_clean = qcType[1] != '0';
    o << "<html lang=\"en\">\n<head>\n<meta http-equiv=\"content-type\" content=\"text/html;charset=utf-8\" />\n"
         "<title>TGSFilter Report</title>\n<style>body{font-family:sans-serif;margin:2em}table{border-collapse:collapse}"
         "td{border:1px solid #999;padding:4px 10px}.plot{width:720px;height:360px;margin:1em 0}</style>\n</head>\n<body>\n"
         "<h1>TGSFilter Report</h1>\n<h2>Summary</h2>\n";
    auto tr = [&](const std::string &a, const std::string &b, const std::string &c) { // genTableTrTd
        o << "<tr>\n" << "    <td>" << a << "</td>\n";
        if (show_raw) o << "    <td>" << b << "</td>\n";
        if (show_clean) o << "    <td>" << c << "</td>\n";
        o << "</tr>\n";
    };
    auto cell = [](const Side &s, int i) { return s.tab[i].empty() ? std::string("0") : s.tab[i]; };
    o << "<table>\n";
    tr("", qcType[1] == '0' ? "Value" : "Before filtering", "After filtering");
    const char *rows[9] = {"Total reads", "Total bases", "GC content (%)", "Min length (bp)", "Max length (bp)",
                           "Mean length (bp)", "Median length (bp)", "N50 length (bp)", "Mean quality"};
    for (int i = 0; i < 9; ++i)
        if (i < 8 || fq) tr(rows[i], cell(raw, i), cell(clean, i));
    o << "</table>\n";
    o << "<div id=\"plots\"></div>\n";
    o << "<script>\n" << "var data = {\n";
    if (show_raw) {
        put(o, "rawLenDis", raw.lenDis);
        put(o, "rawBasesContents", raw.basesContents);
        put(o, "raw5pBasesContents", raw.basesContents5p);
        put(o, "raw3pBasesContents", raw.basesContents3p);
        if (fq) {
            put(o, "rawQualDis", raw.qualDis);
            put(o, "rawReadsQual", raw.readsQual);
            put(o, "raw5pReadsQual", raw.readsQual5p);
            put(o, "raw3pReadsQual", raw.readsQual3p);
        }
    }
    if (show_clean) {
        put(o, "cleanLenDis", clean.lenDis);
        put(o, "cleanBasesContents", clean.basesContents);
        put(o, "clean5pBasesContents", clean.basesContents5p);
        put(o, "clean3pBasesContents", clean.basesContents3p);
        if (fq) {
            put(o, "cleanQualDis", clean.qualDis);
            put(o, "cleanReadsQual", clean.readsQual);
            put(o, "clean5pReadsQual", clean.readsQual5p);
            put(o, "clean3pReadsQual", clean.readsQual3p);
        }
    }
    o << "}\n" << "</script>\n";
    // this repo's own minimal viewer: one canvas per key, drawn by the few lines below (no external library, the
    // report works offline like the reference's, which embeds a 1 MB copy of echarts instead)
    o << "<script>\n"
         "(function () {\n"
         "  var colors = ['#1f77b4', '#d62728', '#2ca02c', '#ff7f0e', '#9467bd', '#8c564b'];\n"
         "  function draw(key, d) {\n"
         "    var box = document.createElement('div'); box.className = 'plot';\n"
         "    var cv = document.createElement('canvas'); cv.width = 720; cv.height = 360; box.appendChild(cv);\n"
         "    document.getElementById('plots').appendChild(box);\n"
         "    var g = cv.getContext('2d'), L = 56, R = 12, T = 30, B = 34, W = cv.width - L - R, H = cv.height - T - B;\n"
         "    var multi = d.y.length && typeof d.y[0] === 'object';\n"
         "    var series = multi ? d.y : [{name: '', data: d.y}];\n"
         "    var n = d.x.length, ymax = 0;\n"
         "    series.forEach(function (s) { s.data.forEach(function (v) { if (v > ymax) ymax = v; }); });\n"
         "    if (ymax <= 0) ymax = 1;\n"
         "    g.font = '13px sans-serif'; g.fillStyle = '#000'; g.fillText(key, L, 18);\n"
         "    g.strokeStyle = '#999'; g.strokeRect(L, T, W, H); g.font = '11px sans-serif';\n"
         "    for (var t = 0; t <= 4; t++) { var yy = T + H - H * t / 4; g.fillText((ymax * t / 4).toPrecision(3), 4, yy + 4);\n"
         "      g.beginPath(); g.moveTo(L, yy); g.lineTo(L + W, yy); g.strokeStyle = '#eee'; g.stroke(); }\n"
         "    for (var t = 0; t <= 6 && n > 0; t++) { var i = Math.min(n - 1, Math.round((n - 1) * t / 6));\n"
         "      g.fillStyle = '#000'; g.fillText(String(d.x[i]), L + (n > 1 ? W * i / (n - 1) : 0) - 8, T + H + 16); }\n"
         "    var bar = /LenDis$/.test(key);\n"
         "    series.forEach(function (s, k) {\n"
         "      g.strokeStyle = g.fillStyle = colors[k % colors.length];\n"
         "      if (bar) { var bw = Math.max(1, W / Math.max(n, 1) - 1);\n"
         "        s.data.forEach(function (v, i) { var h = H * v / ymax; g.fillRect(L + W * i / Math.max(n, 1), T + H - h, bw, h); });\n"
         "      } else { g.beginPath();\n"
         "        s.data.forEach(function (v, i) { var x = L + (n > 1 ? W * i / (n - 1) : 0), y = T + H - H * v / ymax;\n"
         "          if (i) g.lineTo(x, y); else g.moveTo(x, y); }); g.stroke(); }\n"
         "      if (s.name) g.fillText(s.name, L + W - 40 * (series.length - k), 18);\n"
         "    });\n"
         "  }\n"
         "  for (var k in data) draw(k, data[k]);\n"
         "})();\n"
         "</script>\n";
    std::time_t now = std::time(nullptr);
    char buf[80];
    std::strftime(buf, sizeof(buf), "%Y-%m-%d %H:%M:%S", std::localtime(&now));
    o << "<p>Generated by tgsfilter (B200 host of TGSFilter v1.11) at " << buf << "</p>\n</body>\n</html>\n";
}

}  // namespace report
