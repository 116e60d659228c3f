/*
 * ref_harness.c -- TEST INFRASTRUCTURE ONLY (never linked into the product library).
 *
 * Thin driver around the UNMODIFIED reference sources. Nothing from the reference is copied:
 * its translation units are pulled in by path at compile time (REF_SRC is passed by
 * oracle/Makefile and points at /root/reference/src), so this file only builds where the
 * reference tree is mounted. The resulting binaries go to oracle/_ref/ (git-ignored) and travel
 * to the GPU box as prebuilt files.
 *
 * Two flavours are built from this one file:
 *   -DREF_WITH_HW   libvfgs_ref.so     reference hw layer + fw layer + cfg parser  (the CPU oracle
 *                                      proper; also exposes the hw statics for state dumps)
 *   (no define)     libvfgs_fwref.so   reference fw layer + cfg parser only; the ten vfgs_hw.h
 *                                      symbols stay UNDEFINED and bind at load time to whichever
 *                                      hw implementation is already loaded (the CUDA shim) --
 *                                      this is the drop-in demonstration.
 *
 * The cfg parser and its file-scope configuration structs live in the reference CLI source
 * (vfgs_main.c); they are reached with the include trick from SURVEY.md appendix B.
 */
#include <stddef.h>
#include <string.h>
#include <stdio.h>

#define REFH_STR2(x) #x
#define REFH_STR(x) REFH_STR2(x)
#define REFH_SRC(name) REFH_STR(REF_SRC/name)

#ifdef REF_WITH_HW
#include REFH_SRC(vfgs_hw.c)      /* brings the hw statics into this TU */
#undef min
#undef max
#undef round
#endif

#define main refh_cli_main        /* keep the reference CLI's main() out of the way */
#include REFH_SRC(vfgs_main.c)
#undef main

#define REFH_API __attribute__((visibility("default")))

static fgs_sei refh_default_sei;
static int refh_have_default = 0;

/* Mirrors what main() does between argument parsing and the frame loop (vfgs_main.c:739-753,
 * and pop_cfg 635-644): fill the CLI's file-scope parameters, parse one cfg (or keep the
 * built-in default SEI when path is NULL), validate, adapt chroma parameters, apply the gain.
 * enforce_check=0 skips the 4:2:2/4:4:4 rejection so luma-only cfgs can be used there. */
REFH_API int refh_load_cfg(const char* path, int w, int h, int bitdepth, int fmt, unsigned gain,
                           int enforce_check)
{
	if (!refh_have_default) { refh_default_sei = sei; refh_have_default = 1; }
	sei = refh_default_sei;
	memset(&afgs1, 0, sizeof(afgs1));
	width = w; height = h; depth = bitdepth; format = fmt;
	/* vfgs_main.c:739, 752-753: the built-in default SEI is validated, adapted and gained first */
	if (check_cfg() && enforce_check)
		return 2;
	adjust_chroma_cfg();
	apply_gain(gain);
	/* vfgs_main.c:635-644 (pop_cfg): the cfg file is read ON TOP of that struct */
	if (path) {
		if (read_cfg(path))
			return 1;
		if (check_cfg() && enforce_check)
			return 2;
		adjust_chroma_cfg();
		apply_gain(gain);
	}
	return 0;
}

REFH_API int refh_is_afgs1(void) { return afgs1.num_y_points != 0; }

/* Raw bytes of the parsed metadata struct (fgs_sei or fgs_afgs1), for the golden fixtures. */
REFH_API const void* refh_cfg_struct(int* size)
{
	if (afgs1.num_y_points) { *size = (int)sizeof(afgs1); return &afgs1; }
	*size = (int)sizeof(sei);
	return &sei;
}

/* vfgs_main.c:750-751 */
REFH_API void refh_setup_hw(int bitdepth, int fmt)
{
	vfgs_set_depth(bitdepth);
	vfgs_set_chroma_subsampling((fmt < YUV_444) ? 2 : 1, (fmt < YUV_422) ? 2 : 1);
}

/* vfgs_main.c:755-758 with the currently loaded cfg */
REFH_API void refh_init_hw(void)
{
	if (afgs1.num_y_points) vfgs_init_afgs1(&afgs1);
	else                    vfgs_init_sei(&sei);
}

/* Same, from struct bytes captured earlier (the GPU box has no cfg files). */
REFH_API int refh_init_from_bytes(int is_afgs1, const void* bytes, int size)
{
	if (is_afgs1) {
		fgs_afgs1 a;
		if (size != (int)sizeof(a)) return 1;
		memcpy(&a, bytes, sizeof(a));
		vfgs_init_afgs1(&a);
	} else {
		fgs_sei s;
		if (size != (int)sizeof(s)) return 1;
		memcpy(&s, bytes, sizeof(s));
		vfgs_init_sei(&s);
	}
	return 0;
}

/* Frame driver: same walk as vfgs_add_grain() (vfgs_main.c:664-682) but over caller-provided
 * planes and strides (in samples). chroma_every_line = (height == cheight). */
REFH_API void refh_add_grain_frame(void* Yp, void* Up, void* Vp, int w, int h, int stride,
                                   int cstride, int bitdepth, int chroma_every_line)
{
	unsigned char *Y = Yp, *U = Up, *V = Vp;
	int sz = bitdepth > 8 ? 2 : 1;
	for (int y = 0; y < h; y++) {
		vfgs_add_grain_line(Y, U, V, y, w);
		Y += (size_t)stride * sz;
		if ((y & 1) || chroma_every_line) { U += (size_t)cstride * sz; V += (size_t)cstride * sz; }
	}
}

/* n frames back to back, tightly packed planar (Y,U,V per frame), in place. */
REFH_API void refh_add_grain_frames_packed(void* buf, int nframes, int w, int h, int cw, int ch,
                                           int bitdepth)
{
	int sz = bitdepth > 8 ? 2 : 1;
	size_t ysz = (size_t)w * h * sz, csz = (size_t)cw * ch * sz;
	unsigned char* p = buf;
	if ((w & 15) == 0 && (cw & 7) == 0 && (h == ch || (h & 1) == 0)) {
		for (int f = 0; f < nframes; f++, p += ysz + 2 * csz)
			refh_add_grain_frame(p, p + ysz, p + ysz + csz, w, h, w, cw, bitdepth, h == ch);
		return;
	}
	/* The hw layer always works on whole 16-sample blocks (vfgs_hw.c:301), so a picture whose
	 * width is not a multiple of 16 needs the stride padding yuv_alloc gives it (yuv.c:65,74:
	 * strides rounded up to 64 samples, heights to 16 lines); the padding is zeroed here. */
	int stride = (w + 63) & ~63, cstride = (cw + 63) & ~63;
	size_t ypad = (size_t)stride * (h + 16) * sz, cpad = (size_t)cstride * (ch + 16) * sz;
	unsigned char* tmp = calloc(1, ypad + 2 * cpad);
	for (int f = 0; f < nframes; f++, p += ysz + 2 * csz) {
		unsigned char* src[3] = { p, p + ysz, p + ysz + csz };
		unsigned char* dst[3] = { tmp, tmp + ypad, tmp + ypad + cpad };
		memset(tmp, 0, ypad + 2 * cpad);
		for (int c = 0; c < 3; c++)
			for (int y = 0; y < (c ? ch : h); y++)
				memcpy(dst[c] + (size_t)y * (c ? cstride : stride) * sz, src[c] + (size_t)y * (c ? cw : w) * sz, (size_t)(c ? cw : w) * sz);
		refh_add_grain_frame(dst[0], dst[1], dst[2], w, h, stride, cstride, bitdepth, h == ch);
		for (int c = 0; c < 3; c++)
			for (int y = 0; y < (c ? ch : h); y++)
				memcpy(src[c] + (size_t)y * (c ? cw : w) * sz, dst[c] + (size_t)y * (c ? cstride : stride) * sz, (size_t)(c ? cw : w) * sz);
	}
	free(tmp);
}

/* yuv.c:216-258 through its own entry point, on tightly packed planes. */
REFH_API void refh_to_8bit_packed(void* dst8, const void* src16, int nframes, int w, int h, int cw, int ch)
{
	size_t ys = (size_t)w * h, cs = (size_t)cw * ch;
	for (int f = 0; f < nframes; f++) {
		yuv d, s;
		unsigned char* dp = (unsigned char*)dst8 + f * (ys + 2 * cs);
		unsigned char* sp = (unsigned char*)src16 + f * (ys + 2 * cs) * 2;
		d.Y = dp; d.U = dp + ys; d.V = dp + ys + cs;
		s.Y = sp; s.U = sp + 2 * ys; s.V = sp + 2 * (ys + cs);
		d.width = s.width = w; d.height = s.height = h; d.stride = s.stride = w;
		d.cwidth = s.cwidth = cw; d.cheight = s.cheight = ch; d.cstride = s.cstride = cw;
		d.depth = 8; s.depth = 10;
		yuv_to_8bit(&d, &s);
	}
}

#ifdef REF_WITH_HW
/* ---- access to the hw layer's file-scope state (vfgs_hw.c:49-63) ---- */
typedef struct refh_state_s {
	signed char   pattern[2][VFGS_MAX_PATTERNS + 1][64][64];
	unsigned char slut[3][256];
	unsigned char plut[3][256];
	unsigned int  rnd, rnd_up, line_rnd, line_rnd_up;
	int scale_shift, bs, y_min, y_max, c_min, c_max, csubx, csuby;
} refh_state;

REFH_API int refh_state_size(void) { return (int)sizeof(refh_state); }

REFH_API void refh_get_state(refh_state* s)
{
	memcpy(s->pattern, pattern, sizeof(pattern));
	memcpy(s->slut, sLUT, sizeof(sLUT));
	memcpy(s->plut, pLUT, sizeof(pLUT));
	s->rnd = rnd; s->rnd_up = rnd_up; s->line_rnd = line_rnd; s->line_rnd_up = line_rnd_up;
	s->scale_shift = scale_shift; s->bs = bs;
	s->y_min = Y_min; s->y_max = Y_max; s->c_min = C_min; s->c_max = C_max;
	s->csubx = csubx; s->csuby = csuby;
}

/* Back to the power-on values of vfgs_hw.c:49-63. */
REFH_API void refh_reset_hw(void)
{
	memset(pattern, 0, sizeof(pattern));
	memset(sLUT, 0, sizeof(sLUT));
	memset(pLUT, 0, sizeof(pLUT));
	rnd = rnd_up = line_rnd = line_rnd_up = 0xdeadbeef;
	scale_shift = 5 + 6; bs = 0;
	Y_min = 0; Y_max = 255; C_min = 0; C_max = 255;
	csubx = 2; csuby = 2;
	memset(grain, 0, sizeof(grain));
	memset(scale, 0, sizeof(scale));
}

REFH_API void refh_set_raw_rnd(unsigned int v) { rnd = rnd_up = line_rnd = line_rnd_up = v; }

REFH_API unsigned int refh_prng(unsigned int x, unsigned int nsteps)
{
	while (nsteps--) x = prng(x);
	return x;
}

/* out = {sign, ox, oy} before the j/suby line term; uses the currently set csubx/csuby. */
REFH_API void refh_get_offsets(int c, unsigned int val, int out[3])
{
	int s; uint8 x, y;
	if (c == 0)      get_offset_y(val, &s, &x, &y);
	else if (c == 1) get_offset_u(val, &s, &x, &y);
	else             get_offset_v(val, &s, &x, &y);
	out[0] = s; out[1] = x; out[2] = y;
}
#endif
