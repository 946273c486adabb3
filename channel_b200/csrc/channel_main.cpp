// channel_main.cpp - C++ restatement of PROGRAM channel (channel.f90:16-193) on top of the C ABI.
//
// The north star keeps the Fortran driver in charge (INTEGRATION.md shows the iso_c_binding patch); no
// Fortran compiler exists in this image, so this is the compiled host-side mirror of that driver: same
// input (dns.in, optional Dati.cart.out / Runtimedata, optional coriolis.in), same control flow, same
// outputs (Runtimedata, Dati.cart.out, Dati.cart.<n>.out at the dt_field / dt_save cadence of outstats).
// One process = one GPU = npx 1 (the multi-GPU path is driven through the Fortran shim or torchrun).
//
//   channel_b200_run [--dir D] [--bodyforce coriolis|am_f1|am_butterfly] [--convvel] [--device N] [--check-input]
//   (the body force is a compile-time choice in the reference: -Dbodyforce + an #include of one of body_forces/*/*.inc)
//
// --check-input parses dns.in (and Runtimedata, if any) and prints what the run would use, without a GPU.
#include <cerrno>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/channel_b200.h"
#include "../../include/channel_b200_host.h"

struct DnsIn {   // the 12 lines of dns.in (read_dnsin, dnsdata.f90:98-125)
    int nx = 0, ny = 0, nz = 0;
    double alfa0 = 0, beta0 = 0, re = 0, a = 0, ymin = 0, ymax = 0;
    bool CPI = false;
    int CPI_type = 0;
    double gamma = 0, meanpx = 0, meanpz = 0, meanflowx = 0, meanflowz = 0, u0 = 0, uN = 0;
    double deltat = 0, cflmax = 0, time = 0, dt_field = 0, dt_save = 0, t_max = 0;
    bool time_from_restart = true;
    int nstep = 0, npy = 1;
};

static void die(const std::string& msg) {   // the reference STOPs on errors
    fprintf(stderr, "%s\n", msg.c_str());
    exit(1);
}
static void check(int rc, const char* what) {
    if (rc != 0) die(std::string("channel_b200: ") + what + " failed (code " + std::to_string(rc) + "): " + chb_last_error());
}

// one list-directed record: tokens up to a `!` comment, separated by blanks or commas
static std::vector<std::string> record(const std::string& line) {
    std::string body = line.substr(0, line.find('!'));
    for (char& c : body)
        if (c == ',') c = ' ';
    std::istringstream is(body);
    std::vector<std::string> t;
    for (std::string s; is >> s;) t.push_back(s);
    return t;
}
static double fortran_real(std::string s) {   // 1.5d0, 1.5D-3, 12431
    for (char& c : s)
        if (c == 'd' || c == 'D') c = 'e';
    char* end = nullptr;
    const double v = strtod(s.c_str(), &end);
    if (end == s.c_str()) die("dns.in: cannot read a number from '" + s + "'");
    return v;
}
static bool fortran_logical(std::string s) {  // .TRUE. T t .true. / .FALSE. F
    if (!s.empty() && s[0] == '.') s = s.substr(1);
    if (s.empty()) die("dns.in: empty logical");
    const char c = s[0];
    if (c == 'T' || c == 't') return true;
    if (c == 'F' || c == 'f') return false;
    die("dns.in: cannot read a logical from '" + s + "'");
    return false;
}

static DnsIn read_dnsin(const std::string& path) {
    std::ifstream f(path);
    if (!f) die("cannot open " + path);
    std::vector<std::vector<std::string>> rows;
    for (std::string line; std::getline(f, line);) {
        auto t = record(line);
        if (!t.empty()) rows.push_back(t);
    }
    if (rows.size() < 12) die(path + ": expected 12 data lines, found " + std::to_string(rows.size()));
    auto need = [&](size_t r, size_t n) {
        if (rows[r].size() < n) die(path + ": line " + std::to_string(r + 1) + " needs " + std::to_string(n) + " values");
    };
    DnsIn d;
    need(0, 3); d.nx = atoi(rows[0][0].c_str()); d.ny = atoi(rows[0][1].c_str()); d.nz = atoi(rows[0][2].c_str());
    need(1, 2); d.alfa0 = fortran_real(rows[1][0]); d.beta0 = fortran_real(rows[1][1]);
    need(2, 1); d.re = fortran_real(rows[2][0]);                                         // ni = 1/ni, dnsdata.f90:115
    need(3, 3); d.a = fortran_real(rows[3][0]); d.ymin = fortran_real(rows[3][1]); d.ymax = fortran_real(rows[3][2]);
    need(4, 3); d.CPI = fortran_logical(rows[4][0]); d.CPI_type = atoi(rows[4][1].c_str()); d.gamma = fortran_real(rows[4][2]);
    need(5, 2); d.meanpx = fortran_real(rows[5][0]); d.meanpz = fortran_real(rows[5][1]);
    need(6, 2); d.meanflowx = fortran_real(rows[6][0]); d.meanflowz = fortran_real(rows[6][1]);
    need(7, 2); d.u0 = fortran_real(rows[7][0]); d.uN = fortran_real(rows[7][1]);
    need(8, 3); d.deltat = fortran_real(rows[8][0]); d.cflmax = fortran_real(rows[8][1]); d.time = fortran_real(rows[8][2]);
    need(9, 4); d.dt_field = fortran_real(rows[9][0]); d.dt_save = fortran_real(rows[9][1]); d.t_max = fortran_real(rows[9][2]);
    d.time_from_restart = fortran_logical(rows[9][3]);
    need(10, 1); d.nstep = atoi(rows[10][0].c_str());
    need(11, 1); d.npy = atoi(rows[11][0].c_str());
    return d;
}

// ---- Runtimedata (init_memory dnsdata.f90:159-175, get_record :181-218) --------------------------------
// Returns the lines to keep: everything before the first record whose time matches `threshold` within half
// of its own time step (the next WRITE overwrites that record, as after BACKSPACE); sets *deltat to that
// record's time step and *found.  No match: all lines plus an empty one ("Skipping one line and appending").
static std::vector<std::string> get_record(const std::string& path, double threshold, double* deltat, bool* found) {
    std::ifstream f(path);
    std::vector<std::string> keep;
    *found = false;
    for (std::string line; std::getline(f, line);) {
        auto t = record(line);
        if (t.size() >= 11) {
            const double selectime = fortran_real(t[0]), curr_dt = fortran_real(t[10]);
            if (std::fabs(selectime - threshold) < 0.5 * curr_dt) {
                *deltat = curr_dt;
                *found = true;
                return keep;
            }
        }
        keep.push_back(line);
    }
    keep.push_back("");
    return keep;
}

static std::string runtimedata_line(const double* v) {   // WRITE(101,*) of 11 reals, dnsdata.f90:878
    std::string s;
    char buf[64];
    for (int i = 0; i < 11; ++i) {
        snprintf(buf, sizeof(buf), "%s%24.16E", i ? " " : "", v[i]);
        s += buf;
    }
    return s;
}

struct Run {
    DnsIn p;
    chb_handle h = nullptr;
    chb_host_tables tab;
    std::vector<double> y, d0, d1, d2, d4, D0mat;
    double time = 0, time0 = 0, deltat = 0, ni = 0;
    int istep = 0, ifield = 0;
    bool prev_was_close = false, bodyforce = false, convvel = false;
    FILE* rtd = nullptr;
    std::string dir;

    // outstats (dnsdata.f90:853-918)
    void outstats() {
        double cfl, fr[3], corrpx, corrpz, meanpx, meanpz, Ulo[5], Uhi[5], Wlo[5], Whi[5];
        check(chb_get_step_scalars(h, &cfl, fr, &corrpx, &corrpz, &meanpx, &meanpz, Ulo, Uhi, Wlo, Whi), "chb_get_step_scalars");
        const double runtime_global = cfl;
        if (p.cflmax > 0) deltat = p.cflmax / runtime_global;                             // :862
        double dudy0 = 0, dwdy0 = 0, dudyN = 0, dwdyN = 0;
        for (int j = 0; j < 5; ++j) {                                                     // :866-870
            dudy0 += tab.d140[j] * Ulo[j]; dwdy0 += tab.d140[j] * Wlo[j];
            dudyN -= tab.d14n[j] * Uhi[j]; dwdyN -= tab.d14n[j] * Whi[j];
        }
        const double line[11] = {time, dudy0, dudyN, dwdy0, dwdyN, fr[0] + corrpx * fr[2], meanpx + corrpx,
                                 fr[1] + corrpz * fr[2], meanpz + corrpz, runtime_global * deltat, deltat};
        printf("%10.4f   %11.6f   %11.6f   %11.6f   %11.6f   %9.4f   %9.4f   %9.4f   %9.4f   %9.6f   %9.6f   \n", line[0],
               line[1], line[2], line[3], line[4], line[5], line[6], line[7], line[8], line[9], line[10]);   // :876
        fprintf(rtd, "%s\n", runtimedata_line(line).c_str());                                                // :878
        fflush(rtd);
        if (p.dt_save > 0 &&                                                                                 // :882-886
            std::floor((time + 0.5 * deltat) / p.dt_save) > std::floor((time - 0.5 * deltat) / p.dt_save) && istep > 1) {
            printf(" Writing Dati.cart.out at time %g\n", time);
            save("Dati.cart.out", 0);
        }
        if (time + deltat >= (ifield + 1) * p.dt_field) {                                                    // :895-918
            if (prev_was_close || (std::floor((time + 0.5 * deltat) / p.dt_field) > std::floor((time - 0.5 * deltat) / p.dt_field) &&
                                   time > time0)) {
                ++ifield;
                const std::string n = std::to_string(ifield);
                printf(" Writing Dati.cart.%s.out at time %g\n", n.c_str(), time);
                save("Dati.cart." + n + ".out", 0);
                if (bodyforce) {
                    printf(" Writing Force.cart.%s.out at time %g\n", n.c_str(), time);
                    save("Force.cart." + n + ".out", 1);
                }
                if (convvel) {   // dnsdata.f90:908-913
                    printf(" Writing Convvel.cart.%s.out at time %g\n", n.c_str(), time);
                    check(chb_save_convvel_file(h, (dir + "Convvel.cart." + n + ".out").c_str()), "chb_save_convvel_file");
                }
                prev_was_close = false;
            } else {
                prev_was_close = true;
            }
        }
    }
    // snapshots drain in the background while the time loop continues (restart_io.cu)
    void save(const std::string& name, int field) {
        check(chb_save_restart_file(h, (dir + name).c_str(), time, field, 1), "chb_save_restart_file");
    }
};

int main(int argc, char** argv) {
    std::string dir = "./";
    bool coriolis = false, check_input = false, r_convvel = false;
    std::string am;   // "am_f1" | "am_butterfly"
    int device = 0;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        if (a == "--dir" && i + 1 < argc) { dir = argv[++i]; if (dir.back() != '/') dir += '/'; }
        else if (a == "--coriolis") coriolis = true;
        else if (a == "--bodyforce" && i + 1 < argc) {
            const std::string b = argv[++i];
            if (b == "coriolis") coriolis = true;
            else if (b == "am_f1" || b == "am_butterfly") am = b;
            else die("--bodyforce: coriolis, am_f1 or am_butterfly");
        }
        else if (a == "--device" && i + 1 < argc) device = atoi(argv[++i]);
        else if (a == "--check-input") check_input = true;
        else if (a == "--convvel") r_convvel = true;
        else die("usage: channel_b200_run [--dir D] [--bodyforce coriolis|am_f1|am_butterfly] [--convvel] [--device N] [--check-input]");
    }
    Run r;
    r.dir = dir;
    r.p = read_dnsin(dir + "dns.in");
    DnsIn& p = r.p;
    const double deltat_from_dnsin = p.deltat;
    if (p.npy != 1) die("channel_b200 needs npy=1 (full y-columns local)");
    int nxd = 0, nzd = 0;
    if (chb_host_padded_sizes(p.nx, p.nz, &nxd, &nzd)) die("bad nx/nz");
    r.ni = 1.0 / p.re;
    r.time = p.time;
    r.deltat = p.deltat;

    // Runtimedata: continue an existing file from the restart time, or start a new one   (dnsdata.f90:159-175)
    const std::string rtd_path = dir + "Runtimedata";
    bool rtd_exists = false;
    { std::ifstream t(rtd_path); rtd_exists = (bool)t; }

    if (check_input) {
        printf("{\"nx\": %d, \"ny\": %d, \"nz\": %d, \"nxd\": %d, \"nzd\": %d, \"alfa0\": %.17g, \"beta0\": %.17g, \"ni\": %.17g, "
               "\"a\": %.17g, \"ymin\": %.17g, \"ymax\": %.17g, \"CPI\": %d, \"CPI_type\": %d, \"gamma\": %.17g, \"meanpx\": %.17g, "
               "\"meanpz\": %.17g, \"meanflowx\": %.17g, \"meanflowz\": %.17g, \"u0\": %.17g, \"uN\": %.17g, \"deltat\": %.17g, "
               "\"cflmax\": %.17g, \"time\": %.17g, \"dt_field\": %.17g, \"dt_save\": %.17g, \"t_max\": %.17g, "
               "\"time_from_restart\": %d, \"nstep\": %d, \"npy\": %d, \"rtd_exists\": %d",
               p.nx, p.ny, p.nz, nxd, nzd, p.alfa0, p.beta0, r.ni, p.a, p.ymin, p.ymax, (int)p.CPI, p.CPI_type, p.gamma, p.meanpx,
               p.meanpz, p.meanflowx, p.meanflowz, p.u0, p.uN, p.deltat, p.cflmax, p.time, p.dt_field, p.dt_save, p.t_max,
               (int)p.time_from_restart, p.nstep, p.npy, (int)rtd_exists);
        if (rtd_exists) {   // where a run restarted at dns.in's `time` would continue the file
            double dt = 0; bool found = false;
            auto keep = get_record(rtd_path, p.time, &dt, &found);
            printf(", \"rtd_found\": %d, \"rtd_keep_lines\": %zu, \"rtd_deltat\": %.17g", (int)found, keep.size(), dt);
        }
        printf("}\n");
        return 0;
    }

    // init_MPI / init_memory / init_fft / setup_derivatives / setup_boundary_conditions   (channel.f90:34-44)
    check(chb_create(&r.h, p.nx, p.ny, p.nz, nxd, nzd, p.alfa0, p.beta0, r.ni, p.a, p.ymin, p.ymax, 0, 1, nullptr, device), "chb_create");
    const int ny = p.ny;
    r.y.resize(ny + 3); r.d0.resize((size_t)(ny - 1) * 5); r.d1 = r.d0; r.d2 = r.d0; r.d4 = r.d0; r.D0mat.resize((size_t)(ny + 1) * 5);
    r.tab.y = r.y.data(); r.tab.d0 = r.d0.data(); r.tab.d1 = r.d1.data(); r.tab.d2 = r.d2.data(); r.tab.d4 = r.d4.data();
    r.tab.D0mat = r.D0mat.data();
    check(chb_host_setup_tables(ny, p.a, p.ymin, p.ymax, &r.tab), "chb_host_setup_tables");
    check(chb_host_apply_tables(r.h, &r.tab), "chb_set_tables");

    // read_restart_file (dnsdata.f90:677-720)
    {
        double t = 0;
        const int rc = chb_read_restart_file(r.h, (dir + "Dati.cart.out").c_str(), &t);
        if (rc == 0) {
            printf(" Reading from file Dati.cart.out\n");
            r.time = t;
        } else if (rc == 4 && std::string(chb_last_error()).find("cannot open") != std::string::npos) {
            printf(" Generating initial field...\n");                                     // :705-719: laminar Poiseuille in mode (0,0)
            const size_t col = (size_t)ny + 3, ncol = (size_t)(p.nx + 1) * (2 * p.nz + 1);
            std::vector<double> V(3 * ncol * col * 2, 0.0);                               // [c][ix][iz][iy] complex
            const size_t c00 = ((size_t)0 * (2 * p.nz + 1) + p.nz) * col;                 // ix = 0, iz = 0, component 1
            for (int iy = 0; iy < ny + 3; ++iy) V[2 * (c00 + iy)] = 3 * 0.5 * r.y[iy] * (2 - r.y[iy]);
            check(chb_upload_V(r.h, V.data()), "chb_upload_V");
        } else {
            check(rc, "chb_read_restart_file");
        }
    }
    // move cursor to desired record (channel.f90:53-58)
    std::vector<std::string> keep;
    if (p.time_from_restart && rtd_exists) {
        printf(" Found existing Runtimedata...\n");
        bool found = false;
        keep = get_record(rtd_path, r.time, &r.deltat, &found);
        if (found) printf(" In Runtimedata: starting from time %g\n", r.time);
        else printf(" WARNING: no instant of time matching restart file has been found in Runtimedata. Skipping one line and appending.\n");
    } else {
        if (!p.time_from_restart) { r.time = p.time; r.deltat = p.deltat; }              // CALL read_dnsin()
        printf(" Creating new Runtimedata.\n");
    }
    r.rtd = fopen(rtd_path.c_str(), "w");
    if (!r.rtd) die("cannot open " + rtd_path + ": " + strerror(errno));
    for (auto& l : keep) fprintf(r.rtd, "%s\n", l.c_str());

    r.ifield = (int)std::floor((r.time + 0.5 * r.deltat) / p.dt_field);                  // channel.f90:66
    r.time0 = r.time;
    if (p.cflmax == 0) r.deltat = deltat_from_dnsin;                                     // :70-72

    printf(" \n !====================================================!\n !                     D   N   S                      !\n"
           " !====================================================!\n \n");
    printf("   nx =%5d   ny =%5d   nz =%5d\n   nxd =%5d  nzd =%5d\n", p.nx, p.ny, p.nz, nxd, nzd);
    printf("   alfa0 =%11.6f       beta0 =%11.6f   ni =%8.6f\n", p.alfa0, p.beta0, r.ni);
    printf("   meanpx =%11.6f      meanpz =%11.6f\n   meanflowx =%11.6f   meanflowz =%11.6f\n", p.meanpx, p.meanpz, p.meanflowx, p.meanflowz);
    printf("   nsteps =%6d   time_from_restart =%c\n \n", p.nstep, p.time_from_restart ? 'T' : 'F');

    check(chb_set_forcing(r.h, p.meanpx, p.meanpz, p.meanflowx, p.meanflowz, (int)p.CPI, p.CPI_type, p.gamma), "chb_set_forcing");
    check(chb_set_wall_velocity(r.h, p.u0, p.uN), "chb_set_wall_velocity");
    if (coriolis) {   // config_body_force, body_forces/coriolis/coriolis.inc:4-27
        std::ifstream f(dir + "coriolis.in");
        if (!f) die("cannot open " + dir + "coriolis.in");
        double v[3];
        for (int i = 0; i < 3; ++i) {
            std::string line;
            std::getline(f, line);
            auto t = record(line);
            if (t.empty()) die("coriolis.in: expected 3 values");
            v[i] = fortran_real(t[0]);
        }
        const double omega2 = v[0], kz_cutoff = v[1], y_bot = v[2], y_top = p.ymax - y_bot;
        const int iz_thr = std::min(p.nz, (int)std::floor(kz_cutoff / p.beta0));
        printf(" Using bodyforce, Coriolis force\n Ro(tation number) %g\n kz_cutoff, iz_cutoff %g %d\n y_threshold_bottom %g\n", omega2 / 2,
               kz_cutoff, iz_thr, y_bot);
        std::vector<double> my(ny + 3), mz(2 * p.nz + 1);
        for (int i = 0; i < ny + 3; ++i) my[i] = (r.y[i] <= y_bot || r.y[i] >= y_top) ? 1.0 : 0.0;
        for (int iz = -p.nz; iz <= p.nz; ++iz) mz[iz + p.nz] = std::abs(iz) <= iz_thr ? 1.0 : 0.0;
        const double A[9] = {0, -omega2, 0, omega2, 0, 0, 0, 0, 0};                      // F1 = -2Ro v, F2 = +2Ro u (coriolis.inc:29-41)
        check(chb_set_body_force_linear(r.h, 1, A, my.data(), mz.data(), 0), "chb_set_body_force_linear");
        check(chb_set_body_force(r.h), "chb_set_body_force");
        r.bodyforce = true;
    }

    if (!am.empty()) {   // am_f1.inc / am_butterfly.inc with the hard-coded parameters of am_pardec.inc:1
        const double lambdaz_f = 500.0, amp = 1000.0, PI = 3.1415926535897932384626433832795028841971;
        const int iz_f = (int)std::lround((2.0 * PI / lambdaz_f) / (p.beta0 / 1000.0));
        printf(" Using bodyforce, %s\n", am == "am_f1" ? "am_f1 (suppression of superposition)"
                                                       : "am_buttefly (suppression of superposition + small scales)");
        const int nzt = 2 * p.nz + 1;
        std::vector<double> m((size_t)(ny + 3) * nzt, 0.0);
        for (int i = 0; i < ny + 3; ++i) {
            const double yp = (r.y[i] > 1 ? p.ymax - r.y[i] : r.y[i]) * 1000.0;
            for (int iz = -p.nz; iz <= p.nz; ++iz) {
                bool on;
                if (am == "am_f1") {
                    const double lzp = iz == 0 ? 1e10 : 2 * PI / (p.beta0 * std::abs(iz)) * 1000;
                    on = std::abs(iz) <= iz_f && lzp > 2.3 * yp * yp;
                } else {
                    on = std::abs(iz) <= iz_f ? yp <= 60 : yp > 60;
                }
                m[(size_t)i * nzt + iz + p.nz] = on ? 1.0 : 0.0;
            }
        }
        const double A[9] = {-amp, 0, 0, 0, -amp, 0, 0, 0, -amp};
        check(chb_set_body_force_linear_yz(r.h, 1, A, m.data(), 1), "chb_set_body_force_linear_yz");
        check(chb_set_body_force(r.h), "chb_set_body_force");
        r.bodyforce = true;
    }

    if (r_convvel) {   // #define convvel (header.h:44)
        check(chb_set_convvel(r.h, 1), "chb_set_convvel");
        r.convvel = true;
    }

    // Compute CFL, flow rate, CPI (channel.f90:95-115), first Runtimedata line
    if (r.deltat == 0) r.deltat = 1.0;
    check(chb_cfl_prepass(r.h), "chb_cfl_prepass");
    r.outstats();

    static const double RK[3][3] = {{120.0 / 32.0, 2.0, 0.0}, {120.0 / 8.0, 50.0 / 8.0, 34.0 / 8.0}, {120.0 / 20.0, 90.0 / 20.0, 50.0 / 20.0}};
    // bc0(0,0)%u=u0; bcn(0,0)%u=uN is re-assigned every step in the reference (channel.f90:122-124) with the
    // constants of dns.in: set once above (chb_set_wall_velocity), no per-step host round trip
    while (r.time < p.t_max - r.deltat / 2.0 && r.istep < p.nstep) {                      // channel.f90:118
        ++r.istep;
        for (int k = 0; k < 3; ++k) {
            r.time += 2.0 / RK[k][0] * r.deltat;
            if (r.bodyforce) check(chb_set_body_force(r.h), "chb_set_body_force");
            check(chb_buildrhs(r.h, RK[k], r.deltat, k == 2), "chb_buildrhs");
            check(chb_linsolve(r.h, RK[k][0] / r.deltat), "chb_linsolve");
        }
        r.outstats();
    }
    printf(" End of time/iterations loop: writing restart file at time %g\n", r.time);
    r.save("Dati.cart.out", 0);                                                          // channel.f90:181
    check(chb_restart_wait(r.h), "chb_restart_wait");
    fclose(r.rtd);
    check(chb_destroy(r.h), "chb_destroy");
    return 0;
}
