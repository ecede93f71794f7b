/* Minimal stand-in for <gtk/gtk.h> + glib, ONLY so that the reference's
 * src/stacking/stacking.c and src/algos/statistics.c compile unmodified for
 * the test oracle (oracle/_ref).  Test infrastructure, never shipped.
 * Types are opaque; every GTK entry point the two files mention is declared
 * here and defined as a no-op in ../gui_dummies.c. */
#ifndef ORACLE_STUB_GTK_H
#define ORACLE_STUB_GTK_H
#include <stddef.h>
#include <stdint.h>
#include <assert.h>
#include <limits.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>
#include <sys/types.h>
#include <unistd.h>

typedef int gboolean;
typedef int gint;
typedef unsigned int guint;
typedef char gchar;
typedef unsigned char guchar;
typedef double gdouble;
typedef float gfloat;
typedef long glong;
typedef unsigned long gulong;
typedef size_t gsize;
typedef void *gpointer;
typedef const void *gconstpointer;
typedef uint64_t guint64;
typedef int64_t gint64;
typedef uint32_t guint32;
typedef uint16_t guint16;
typedef uint8_t guint8;
typedef gboolean (*GSourceFunc)(gpointer);
#ifndef TRUE
#define TRUE 1
#endif
#ifndef FALSE
#define FALSE 0
#endif
#define g_assert(x) assert(x)
#define GINT_TO_POINTER(i) ((gpointer)(long)(i))
#define GPOINTER_TO_INT(p) ((gint)(long)(p))
#define G_GNUC_UNUSED __attribute__((unused))

typedef struct _GThread GThread;
typedef struct { void *p; unsigned i[2]; } GMutex;
typedef struct _GSList { void *data; struct _GSList *next; } GSList;
typedef struct _GList { void *data; struct _GList *next, *prev; } GList;
typedef struct _GdkRGBA { double red, green, blue, alpha; } GdkRGBA;
typedef struct _cairo_surface cairo_surface_t;
typedef struct _cairo cairo_t;
typedef struct _GdkEvent GdkEvent;
typedef struct _GdkEventButton GdkEventButton;
typedef struct _GdkEventMotion GdkEventMotion;
typedef struct _GdkEventKey GdkEventKey;
typedef struct _GdkEventCrossing GdkEventCrossing;
typedef struct _GdkEventScroll GdkEventScroll;
typedef struct _GdkEventConfigure GdkEventConfigure;
typedef struct _GdkPixbuf GdkPixbuf;
typedef struct _GObject GObject;
typedef struct _GFile GFile;

#define ORACLE_GTK_TYPE(T) typedef struct _##T T
ORACLE_GTK_TYPE(GtkWidget); ORACLE_GTK_TYPE(GtkBuilder); ORACLE_GTK_TYPE(GtkAdjustment);
ORACLE_GTK_TYPE(GtkComboBox); ORACLE_GTK_TYPE(GtkComboBoxText); ORACLE_GTK_TYPE(GtkEntry);
ORACLE_GTK_TYPE(GtkToggleButton); ORACLE_GTK_TYPE(GtkSpinButton); ORACLE_GTK_TYPE(GtkLabel);
ORACLE_GTK_TYPE(GtkRange); ORACLE_GTK_TYPE(GtkButton); ORACLE_GTK_TYPE(GtkWindow);
ORACLE_GTK_TYPE(GtkTreeView); ORACLE_GTK_TYPE(GtkTreeSelection); ORACLE_GTK_TYPE(GtkTreeModel);
ORACLE_GTK_TYPE(GtkTreePath); ORACLE_GTK_TYPE(GtkTreeViewColumn); ORACLE_GTK_TYPE(GtkListStore);
ORACLE_GTK_TYPE(GtkCellRenderer); ORACLE_GTK_TYPE(GtkCellRendererToggle); ORACLE_GTK_TYPE(GtkCellRendererText);
ORACLE_GTK_TYPE(GtkMenuItem); ORACLE_GTK_TYPE(GtkCheckMenuItem); ORACLE_GTK_TYPE(GtkRadioMenuItem);
ORACLE_GTK_TYPE(GtkMenu); ORACLE_GTK_TYPE(GtkToolButton); ORACLE_GTK_TYPE(GtkToggleToolButton);
ORACLE_GTK_TYPE(GtkRadioButton); ORACLE_GTK_TYPE(GtkCheckButton); ORACLE_GTK_TYPE(GtkFileChooser);
ORACLE_GTK_TYPE(GtkFileChooserButton); ORACLE_GTK_TYPE(GtkFileFilter); ORACLE_GTK_TYPE(GtkNotebook);
ORACLE_GTK_TYPE(GtkTextView); ORACLE_GTK_TYPE(GtkTextBuffer); ORACLE_GTK_TYPE(GtkEditable);
ORACLE_GTK_TYPE(GtkDialog); ORACLE_GTK_TYPE(GtkScale); ORACLE_GTK_TYPE(GtkColorButton);
ORACLE_GTK_TYPE(GtkColorChooser); ORACLE_GTK_TYPE(GtkExpander); ORACLE_GTK_TYPE(GtkStatusbar);
ORACLE_GTK_TYPE(GtkProgressBar); ORACLE_GTK_TYPE(GtkDrawingArea); ORACLE_GTK_TYPE(GtkAboutDialog);
ORACLE_GTK_TYPE(GtkScrolledWindow); ORACLE_GTK_TYPE(GtkPaned); ORACLE_GTK_TYPE(GtkImage);
ORACLE_GTK_TYPE(GtkSwitch); ORACLE_GTK_TYPE(GtkGrid); ORACLE_GTK_TYPE(GtkBox);
ORACLE_GTK_TYPE(GtkEntryBuffer); ORACLE_GTK_TYPE(GtkTreeIter); ORACLE_GTK_TYPE(GtkTextIter);
ORACLE_GTK_TYPE(GtkAllocation); ORACLE_GTK_TYPE(GtkApplication); ORACLE_GTK_TYPE(GtkCssProvider);
ORACLE_GTK_TYPE(GtkIconSize); ORACLE_GTK_TYPE(GtkSelectionData); ORACLE_GTK_TYPE(GdkDragContext);
ORACLE_GTK_TYPE(GtkTooltip); ORACLE_GTK_TYPE(GtkSpinner); ORACLE_GTK_TYPE(GtkFrame);
ORACLE_GTK_TYPE(GtkContainer); ORACLE_GTK_TYPE(GtkMenuToolButton); ORACLE_GTK_TYPE(GtkStack);
typedef int GtkScrollType;
typedef int GtkStateFlags;

/* cast "macros" */
#define GTK_COMBO_BOX(x) ((GtkComboBox *)(x))
#define GTK_COMBO_BOX_TEXT(x) ((GtkComboBoxText *)(x))
#define GTK_ENTRY(x) ((GtkEntry *)(x))
#define GTK_TOGGLE_BUTTON(x) ((GtkToggleButton *)(x))
#define GTK_SPIN_BUTTON(x) ((GtkSpinButton *)(x))
#define GTK_LABEL(x) ((GtkLabel *)(x))
#define GTK_WIDGET(x) ((GtkWidget *)(x))
#define GTK_RANGE(x) ((GtkRange *)(x))
#define GTK_ADJUSTMENT(x) ((GtkAdjustment *)(x))
#define GTK_WINDOW(x) ((GtkWindow *)(x))
#define GTK_BUTTON(x) ((GtkButton *)(x))
#define GTK_NOTEBOOK(x) ((GtkNotebook *)(x))
#define GTK_TREE_VIEW(x) ((GtkTreeView *)(x))
#define GTK_TREE_MODEL(x) ((GtkTreeModel *)(x))
#define GTK_LIST_STORE(x) ((GtkListStore *)(x))
#define GTK_CHECK_MENU_ITEM(x) ((GtkCheckMenuItem *)(x))
#define GTK_EXPANDER(x) ((GtkExpander *)(x))
#define GTK_FILE_CHOOSER(x) ((GtkFileChooser *)(x))
#define GTK_DIALOG(x) ((GtkDialog *)(x))
#define GTK_TEXT_VIEW(x) ((GtkTextView *)(x))
#define G_OBJECT(x) ((GObject *)(x))

GObject *gtk_builder_get_object(GtkBuilder *b, const gchar *name);
void gtk_combo_box_set_active(GtkComboBox *c, gint i);
gint gtk_combo_box_get_active(GtkComboBox *c);
gdouble gtk_spin_button_get_value(GtkSpinButton *s);
void gtk_spin_button_set_value(GtkSpinButton *s, gdouble v);
void gtk_spin_button_set_range(GtkSpinButton *s, gdouble lo, gdouble hi);
void gtk_spin_button_set_increments(GtkSpinButton *s, gdouble a, gdouble b);
gint gtk_spin_button_get_value_as_int(GtkSpinButton *s);
gboolean gtk_toggle_button_get_active(GtkToggleButton *t);
void gtk_toggle_button_set_active(GtkToggleButton *t, gboolean b);
const gchar *gtk_entry_get_text(GtkEntry *e);
void gtk_entry_set_text(GtkEntry *e, const gchar *t);
void gtk_label_set_text(GtkLabel *l, const gchar *t);
void gtk_widget_set_sensitive(GtkWidget *w, gboolean s);
void gtk_widget_set_visible(GtkWidget *w, gboolean s);
gboolean gtk_widget_get_visible(GtkWidget *w);
void gtk_widget_show(GtkWidget *w);
void gtk_widget_hide(GtkWidget *w);
void gtk_widget_set_tooltip_text(GtkWidget *w, const gchar *t);
gdouble gtk_range_get_value(GtkRange *r);
void gtk_range_set_value(GtkRange *r, gdouble v);
gdouble gtk_adjustment_get_value(GtkAdjustment *a);
void gtk_adjustment_set_value(GtkAdjustment *a, gdouble v);
void gtk_adjustment_set_upper(GtkAdjustment *a, gdouble v);
void gtk_notebook_set_current_page(GtkNotebook *n, gint p);
guint gdk_threads_add_idle(GSourceFunc f, gpointer data);
gint g_snprintf(gchar *s, gulong n, const gchar *fmt, ...);
gchar *g_strdup_printf(const gchar *fmt, ...);
void g_free(gpointer p);
#endif
