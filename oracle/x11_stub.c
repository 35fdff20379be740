/*
 * x11_stub.c -- just enough of libX11 / libXext for Mesa's xlib GLX software
 * driver (llvmpipe) to create a context and render to an FBO with NO X server.
 *
 * TEST INFRASTRUCTURE ONLY.  Built twice by oracle/Makefile, as
 * oracle/_ref/libX11.so.6 and oracle/_ref/libXext.so.6 (soname set), so that the
 * Mesa 18.1.9 libGL.so.1 bundled with Nsight Compute can be dlopen'ed; used only
 * by oracle/glsl_ref.c to run the reference's UNMODIFIED GLSL headless
 * (SURVEY.md Appendix B).  Everything is a no-op except what Mesa needs to
 * believe it has a 24-bit TrueColor screen.
 *
 * No X headers exist in this image: the few Xlib structs are declared here with
 * the public Xlib.h layout (x86-64).
 */
#include <stdlib.h>
#include <string.h>

typedef unsigned long XID;
typedef char*         XPointer;

typedef struct
{
    void*         ext_data;
    XID           visualid;
    int           c_class;
    unsigned long red_mask, green_mask, blue_mask;
    int           bits_per_rgb;
    int           map_entries;
} Visual;

typedef struct
{
    int     depth;
    int     nvisuals;
    Visual* visuals;
} Depth;

typedef struct
{
    void*         ext_data;
    void*         display;
    XID           root;
    int           width, height;
    int           mwidth, mheight;
    int           ndepths;
    Depth*        depths;
    int           root_depth;
    Visual*       root_visual;
    void*         default_gc;
    XID           cmap;
    unsigned long white_pixel;
    unsigned long black_pixel;
    int           max_maps, min_maps;
    int           backing_store;
    int           save_unders;
    long          root_input_mask;
} Screen;

typedef struct
{
    void* ext_data;
    int   depth;
    int   bits_per_pixel;
    int   scanline_pad;
} ScreenFormat;

/* public prefix of Display (_XPrivDisplay) followed by zero padding that covers
 * the private tail Mesa touches (ext_procs at byte 0x140) */
typedef struct
{
    void*         ext_data;
    void*         private1;
    int           fd;
    int           private2;
    int           proto_major_version;
    int           proto_minor_version;
    char*         vendor;
    XID           private3, private4, private5;
    int           private6;
    void*         resource_alloc;
    int           byte_order;
    int           bitmap_unit;
    int           bitmap_pad;
    int           bitmap_bit_order;
    int           nformats;
    ScreenFormat* pixmap_format;
    int           private8;
    int           release;
    void *        private9, *private10;
    int           qlen;
    unsigned long last_request_read;
    unsigned long request;
    XPointer      private11, private12, private13, private14;
    unsigned      max_request_size;
    void*         db;
    void*         private15;
    char*         display_name;
    int           default_screen;
    int           nscreens;
    Screen*       screens;
    unsigned long motion_buffer;
    unsigned long private16;
    int           min_keycode;
    int           max_keycode;
    XPointer      private17, private18;
    int           private19;
    char*         xdefaults;
    char          tail[8192];
} FakeDisplay;

typedef struct
{
    Visual*       visual;
    XID           visualid;
    int           screen;
    int           depth;
    int           c_class;
    unsigned long red_mask, green_mask, blue_mask;
    int           colormap_size;
    int           bits_per_rgb;
} XVisualInfo;

typedef struct _XImage
{
    int           width, height;
    int           xoffset;
    int           format;
    char*         data;
    int           byte_order;
    int           bitmap_unit;
    int           bitmap_bit_order;
    int           bitmap_pad;
    int           depth;
    int           bytes_per_line;
    int           bits_per_pixel;
    unsigned long red_mask, green_mask, blue_mask;
    XPointer      obdata;
    struct
    {
        struct _XImage* (*create_image)(void);
        int (*destroy_image)(struct _XImage*);
        unsigned long (*get_pixel)(struct _XImage*, int, int);
        int (*put_pixel)(struct _XImage*, int, int, unsigned long);
        struct _XImage* (*sub_image)(struct _XImage*, int, int, unsigned int, unsigned int);
        int (*add_pixel)(struct _XImage*, long);
    } f;
} XImage;

typedef struct
{
    int           x, y;
    int           width, height;
    int           border_width;
    int           depth;
    Visual*       visual;
    XID           root;
    int           c_class;
    int           bit_gravity;
    int           win_gravity;
    int           backing_store;
    unsigned long backing_planes;
    unsigned long backing_pixel;
    int           save_under;
    XID           colormap;
    int           map_installed;
    int           map_state;
    long          all_event_masks;
    long          your_event_mask;
    long          do_not_propagate_mask;
    int           override_redirect;
    Screen*       screen;
} XWindowAttributes;

#define TrueColor 4
#define VisualIDMask 0x1
#define VisualScreenMask 0x2
#define VisualDepthMask 0x4
#define VisualClassMask 0x8

static Visual       g_visual = {0, 0x21, TrueColor, 0xff0000, 0x00ff00, 0x0000ff, 8, 256};
static Depth        g_depth  = {24, 1, &g_visual};
static Screen       g_screen;
static ScreenFormat g_format = {0, 24, 32, 32};
static FakeDisplay  g_display;
static int          g_win_w = 64, g_win_h = 64;

/* entry point for the harness: there is no XOpenDisplay import in Mesa */
void* qb_stub_open_display(int win_w, int win_h)
{
    memset(&g_display, 0, sizeof(g_display));
    memset(&g_screen, 0, sizeof(g_screen));
    g_win_w              = win_w;
    g_win_h              = win_h;
    g_screen.display     = &g_display;
    g_screen.root        = 1;
    g_screen.width       = 4096;
    g_screen.height      = 4096;
    g_screen.mwidth      = 1000;
    g_screen.mheight     = 1000;
    g_screen.ndepths     = 1;
    g_screen.depths      = &g_depth;
    g_screen.root_depth  = 24;
    g_screen.root_visual = &g_visual;
    g_screen.default_gc  = calloc(1, 256);
    g_screen.cmap        = 3;
    g_screen.white_pixel = 0xffffff;

    g_display.fd                  = -1;
    g_display.proto_major_version = 11;
    g_display.vendor              = "qb-stub";
    g_display.byte_order          = 0; /* LSBFirst */
    g_display.bitmap_unit         = 32;
    g_display.bitmap_pad          = 32;
    g_display.bitmap_bit_order    = 0;
    g_display.nformats            = 1;
    g_display.pixmap_format       = &g_format;
    g_display.release             = 1;
    g_display.display_name        = ":stub";
    g_display.default_screen      = 0;
    g_display.nscreens            = 1;
    g_display.screens             = &g_screen;
    return &g_display;
}

/* XExtCodes* XAddExtension(Display*): Mesa afterwards writes into the private
 * _XExtension record that owns the codes (found at dpy->ext_procs, byte 0x140) */
void* XAddExtension(void* dpy)
{
    char* block = calloc(1, 512);
    *(void**) ((char*) dpy + 0x140) = block;
    int* codes                      = (int*) (block + 8);
    codes[0]                        = 1;   /* extension number */
    codes[1]                        = 128; /* major opcode */
    return codes;
}

int XQueryExtension(void* dpy, const char* name, int* a, int* b, int* c) { return 0; }

XVisualInfo* XGetVisualInfo(void* dpy, long mask, XVisualInfo* tmpl, int* nitems)
{
    *nitems = 0;
    if ((mask & VisualDepthMask) && tmpl->depth != 24) return NULL;
    if ((mask & VisualClassMask) && tmpl->c_class != TrueColor) return NULL;
    if ((mask & VisualIDMask) && tmpl->visualid != g_visual.visualid) return NULL;
    if ((mask & VisualScreenMask) && tmpl->screen != 0) return NULL;
    XVisualInfo* v   = calloc(1, sizeof(*v));
    v->visual        = &g_visual;
    v->visualid      = g_visual.visualid;
    v->screen        = 0;
    v->depth         = 24;
    v->c_class       = TrueColor;
    v->red_mask      = g_visual.red_mask;
    v->green_mask    = g_visual.green_mask;
    v->blue_mask     = g_visual.blue_mask;
    v->colormap_size = 256;
    v->bits_per_rgb  = 8;
    *nitems          = 1;
    return v;
}

static int destroy_image(XImage* img)
{
    if (img)
    {
        free(img->data);
        free(img);
    }
    return 1;
}
static unsigned long get_pixel(XImage* img, int x, int y) { return 0; }
static int           put_pixel(XImage* img, int x, int y, unsigned long p) { return 1; }

XImage* XCreateImage(void* dpy, Visual* visual, unsigned int depth, int format, int offset, char* data,
                     unsigned int width, unsigned int height, int bitmap_pad, int bytes_per_line)
{
    XImage* img           = calloc(1, sizeof(*img));
    img->width            = (int) width;
    img->height           = (int) height;
    img->xoffset          = offset;
    img->format           = format;
    img->data             = data;
    img->byte_order       = 0;
    img->bitmap_unit      = 32;
    img->bitmap_bit_order = 0;
    img->bitmap_pad       = bitmap_pad ? bitmap_pad : 32;
    img->depth            = (int) depth;
    img->bits_per_pixel   = 32;
    img->bytes_per_line   = bytes_per_line ? bytes_per_line : (int) width * 4;
    img->red_mask         = g_visual.red_mask;
    img->green_mask       = g_visual.green_mask;
    img->blue_mask        = g_visual.blue_mask;
    img->f.destroy_image  = destroy_image;
    img->f.get_pixel      = get_pixel;
    img->f.put_pixel      = put_pixel;
    return img;
}

int XGetGeometry(void* dpy, XID d, XID* root, int* x, int* y, unsigned int* w, unsigned int* h, unsigned int* bw,
                 unsigned int* depth)
{
    if (root) *root = 1;
    if (x) *x = 0;
    if (y) *y = 0;
    if (w) *w = (unsigned) g_win_w;
    if (h) *h = (unsigned) g_win_h;
    if (bw) *bw = 0;
    if (depth) *depth = 24;
    return 1;
}

int XGetWindowAttributes(void* dpy, XID w, XWindowAttributes* a)
{
    memset(a, 0, sizeof(*a));
    a->width     = g_win_w;
    a->height    = g_win_h;
    a->depth     = 24;
    a->visual    = &g_visual;
    a->root      = 1;
    a->c_class   = 1; /* InputOutput */
    a->colormap  = 3;
    a->map_state = 2; /* IsViewable */
    a->screen    = &g_screen;
    return 1;
}

XID   XCreateColormap(void* dpy, XID w, Visual* v, int alloc) { return 3; }
void* XCreateGC(void* dpy, XID d, unsigned long mask, void* values) { return calloc(1, 256); }
XID   XCreatePixmap(void* dpy, XID d, unsigned int w, unsigned int h, unsigned int depth) { return 5; }
int   XDrawString16(void* dpy, XID d, void* gc, int x, int y, void* s, int n) { return 0; }
int   XFillRectangle(void* dpy, XID d, void* gc, int x, int y, unsigned int w, unsigned int h) { return 0; }
int   XFlush(void* dpy) { return 0; }
int   XFree(void* p)
{
    free(p);
    return 1;
}
int   XFreeFontInfo(char** names, void* info, int n) { return 0; }
int   XFreeGC(void* dpy, void* gc) { return 0; }
int   XFreePixmap(void* dpy, XID p) { return 0; }
void* XGetImage(void* dpy, XID d, int x, int y, unsigned int w, unsigned int h, unsigned long mask, int fmt)
{
    return NULL;
}
int   XPutImage(void* dpy, XID d, void* gc, XImage* img, int sx, int sy, int dx, int dy, unsigned int w,
                unsigned int h)
{
    return 0;
}
void* XQueryFont(void* dpy, XID id) { return NULL; }
void* XSetErrorHandler(void* h) { return NULL; }
int   XSetForeground(void* dpy, void* gc, unsigned long fg) { return 0; }
int   XSetFunction(void* dpy, void* gc, int f) { return 0; }
int   XShmAttach(void* dpy, void* info) { return 0; }
void* XShmCreateImage(void* dpy, Visual* v, unsigned int depth, int fmt, char* data, void* info, unsigned int w,
                      unsigned int h)
{
    return NULL;
}
int   XShmPutImage(void* dpy, XID d, void* gc, XImage* img, int sx, int sy, int dx, int dy, unsigned int w,
                   unsigned int h, int send_event)
{
    return 0;
}
int   XSync(void* dpy, int discard) { return 0; }
void* XSynchronize(void* dpy, int onoff) { return NULL; }

void (*_XLockMutex_fn)(void*)   = NULL;
void (*_XUnlockMutex_fn)(void*) = NULL;
void* _Xglobal_lock             = NULL;
