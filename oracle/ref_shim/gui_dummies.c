/* oracle/ref_shim/gui_dummies.c -- TEST INFRASTRUCTURE ONLY.
 *
 * The reference's stacking.c also holds the GTK glue of the stacking tab
 * (start_stacking, end_stacking, update_stack_interface ...).  None of it is
 * on the hot path and none of it is called by the oracle, but the object file
 * references these symbols, so they are given no-op bodies here.  This file
 * deliberately includes no reference header (the K&R-style bodies would clash
 * with the real prototypes). */
#include <stddef.h>

#define ORACLE_NOOP(name) long name() { return 0; }

ORACLE_NOOP(adjust_cutoff_from_updated_gfit)
ORACLE_NOOP(clear_stars_list)
ORACLE_NOOP(control_window_switch_to_tab)
ORACLE_NOOP(display_filename)
ORACLE_NOOP(gdk_threads_add_idle)
ORACLE_NOOP(gtk_adjustment_get_value)
ORACLE_NOOP(gtk_builder_get_object)
ORACLE_NOOP(gtk_combo_box_get_active)
ORACLE_NOOP(gtk_combo_box_set_active)
ORACLE_NOOP(gtk_entry_get_text)
ORACLE_NOOP(gtk_label_set_text)
ORACLE_NOOP(gtk_notebook_set_current_page)
ORACLE_NOOP(gtk_spin_button_get_adjustment)
ORACLE_NOOP(gtk_spin_button_get_value)
ORACLE_NOOP(gtk_spin_button_set_range)
ORACLE_NOOP(gtk_spin_button_set_value)
ORACLE_NOOP(gtk_toggle_button_get_active)
ORACLE_NOOP(gtk_widget_set_sensitive)
ORACLE_NOOP(initialize_display_mode)
ORACLE_NOOP(lookup_widget)
ORACLE_NOOP(noise)
ORACLE_NOOP(redraw)
ORACLE_NOOP(redraw_previews)
ORACLE_NOOP(savefits)
ORACLE_NOOP(sequence_list_change_current)
ORACLE_NOOP(set_cursor_waiting)
ORACLE_NOOP(set_cutoff_sliders_max_values)
ORACLE_NOOP(set_display_mode)
ORACLE_NOOP(set_sliders_value_to_gfit)
ORACLE_NOOP(sliders_mode_set_state)
ORACLE_NOOP(show_dialog)
ORACLE_NOOP(show_time)
ORACLE_NOOP(start_in_new_thread)
ORACLE_NOOP(stop_processing_thread)
ORACLE_NOOP(update_MenuItem)
ORACLE_NOOP(writeinitfile)
ORACLE_NOOP(writeseqfile)
ORACLE_NOOP(update_used_memory)
ORACLE_NOOP(sequence_is_loaded)

void *builder = NULL;
